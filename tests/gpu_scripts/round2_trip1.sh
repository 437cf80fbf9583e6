#!/bin/bash
# Round-2 GPU trip 1: parity of the refactored library (one graph per plan, pruned variants), the new bench line with the
# other BASELINE.json configs / library baseline / batch-1 block, batch-1 per-layer profile, ncu launch lists and full
# captures (normal batch 8; 64-channel stacked pair kernel), compute-sanitizer.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
LSPG_PER_LAYER=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 1 > gpurun_out/b1_layers.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 212 --csv --log-file gpurun_out/launches_b1.csv \
    python tests/gpu_bringup.py time large A parity 512 1 > gpurun_out/ncu_list_b1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 420 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/ncu_list.log 2>&1
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_pair_kernel' -s 8 -c 3 -o gpurun_out/prof_normal_b8 -f \
    python bench.py --variant normal --batch 8 --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_normal_b8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_pair_kernel<\(int\)64' -s 9 -c 2 -o gpurun_out/prof_pair64 -f \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_pair64.log 2>&1
bash tests/gpu_scripts/sanitize.sh 240
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_ref.json
head -12 gpurun_out/b1_layers.log; ls -la gpurun_out | head -50
