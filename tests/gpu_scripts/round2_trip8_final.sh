#!/bin/bash
# Round-2 final single-GPU trip: the GPU suite, smoke, the bench line as the driver runs it (default flags + the driver's short
# form), the reference arm, the ncu launch list of the bench command and one ncu --set full capture of the widest kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_driver_form.json 2> gpurun_out/bench_driver_form.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 320 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/ncu_list.log 2>&1
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_pair_kernel<\(int\)256' -s 8 -c 2 -o gpurun_out/prof_pair256 -f \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_pair256.log 2>&1
tail -14 gpurun_out/pytest_gpu.log; tail -4 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench.json; tail -2 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_driver_form.json; cut -c1-300 gpurun_out/bench_ref.json
