#!/bin/bash
# Round-2 trip 7 (1 GPU): frames-per-step sweep of the final kernels, full GPU suite, smoke, bench + launch list + ncu full capture.
mkdir -p gpurun_out
for B in 32 48 64 96 128; do timeout 200 python tests/gpu_bringup.py time large A parity 512 $B >> gpurun_out/batch_sweep.log 2>&1; done
timeout 200 python tests/gpu_bringup.py time large A parity 512 32 >> gpurun_out/batch_sweep.log 2>&1
timeout 200 python tests/gpu_bringup.py time large A parity 512 64 >> gpurun_out/batch_sweep.log 2>&1
timeout 200 python tests/gpu_bringup.py time normal A parity 512 64 >> gpurun_out/batch_sweep.log 2>&1
grep "ms/forward" gpurun_out/batch_sweep.log
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --batch 64 > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "bench rc=$?" >> gpurun_out/bench_b64.err
timeout 300 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_pair_kernel<\(int\)256' -s 8 -c 2 -o gpurun_out/prof_pair256_fp16 -f \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_pair256.log 2>&1
tail -12 gpurun_out/pytest_gpu.log; tail -4 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_b64.json; tail -2 gpurun_out/bench_b64.err
