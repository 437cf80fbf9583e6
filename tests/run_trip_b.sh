#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
{
T="timeout 300 python tests/gpu_bringup.py"
echo "## stack final";              $T final large A parity 512 16
echo "## stack final normal B";     $T final normal B parity 512 4
echo "## stack";                    $T time large A parity 512 16
echo "## nostack"; LSPG_NO_PAIR_STACK=1 $T time large A parity 512 16
echo "## stack";                    $T time large A parity 512 16
echo "## nostack"; LSPG_NO_PAIR_STACK=1 $T time large A parity 512 16
echo "## stack per-layer"; LSPG_PER_LAYER=1 $T time large A parity 512 16
} > gpurun_out/trip_b.log 2>&1
LSPG_TRACE_SKIP=20 LSPG_TRACE_LAYERS=1,2,70,6,12 LSPG_TRACE_CTAS=0,1 timeout 300 python tests/gpu_trace.py large parity 16 > gpurun_out/trace5.log 2>&1
LSPG_TRACE_LAYERS=21 LSPG_TRACE_CTAS=0,1,64 timeout 300 python tests/gpu_trace.py large parity 16 >> gpurun_out/trace5.log 2>&1
rm -f gpurun_out/prof_pair64.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:conv_pair_kernel<64' -s 18 -c 2 -o gpurun_out/prof_pair64 -f \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_pair64.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; grep -E "^##|^large|^normal|max" gpurun_out/trip_b.log | head -40; tail -4 gpurun_out/ncu_pair64.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
