#!/bin/bash
# One GPU trip: parity tests, smoke, bench (both arms), launch list + ncu full captures, library baseline.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python tests/gpu_library_baseline.py large 8 > gpurun_out/cudnn_baseline.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 420 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/ncu_list.log 2>&1
rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_pair_kernel<\(int\)256' -s 16 -c 2 -o gpurun_out/prof_pair256 -f \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_pair256.log 2>&1
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:conv_pair_kernel<\(int\)128' -s 36 -c 2 -o gpurun_out/prof_pair128 -f \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_pair128.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ref.json; cat gpurun_out/cudnn_baseline.log | tail -4; ls -la gpurun_out
