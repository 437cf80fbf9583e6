#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
{
T="timeout 300 python tests/gpu_bringup.py"
echo "## base";                     $T time large A parity 512 16
echo "## rotate";   LSPG_ROTATE=1   $T time large A parity 512 16
echo "## small";    LSPG_SMALL_BN128=1 LSPG_SPLIT_FLOOR=1 $T time large A parity 512 16
echo "## floor";    LSPG_SPLIT_FLOOR=1 $T time large A parity 512 16
echo "## all";      LSPG_ROTATE=1 LSPG_SMALL_BN128=1 LSPG_SPLIT_FLOOR=1 $T time large A parity 512 16
echo "## all final"; LSPG_ROTATE=1 LSPG_SMALL_BN128=1 LSPG_SPLIT_FLOOR=1 $T final large A parity 512 16
echo "## all final normal B"; LSPG_ROTATE=1 LSPG_SMALL_BN128=1 LSPG_SPLIT_FLOOR=1 $T final normal B parity 512 4
echo "## base B1";                  $T time large A parity 512 1
echo "## all B1";   LSPG_ROTATE=1 LSPG_SMALL_BN128=1 LSPG_SPLIT_FLOOR=1 $T time large A parity 512 1
echo "## base fast";                $T time large A fast 512 16
echo "## all per-layer"; LSPG_PER_LAYER=1 LSPG_ROTATE=1 LSPG_SMALL_BN128=1 LSPG_SPLIT_FLOOR=1 $T time large A parity 512 16
} > gpurun_out/trip_a.log 2>&1
LSPG_TRACE_LAYERS=21,26,2 LSPG_TRACE_CTAS=0,1 timeout 300 python tests/gpu_trace.py large parity 16 > gpurun_out/trace4.log 2>&1
LSPG_TRACE_SKIP=20 LSPG_TRACE_LAYERS=2 LSPG_TRACE_CTAS=0,1 timeout 300 python tests/gpu_trace.py large parity 16 >> gpurun_out/trace4.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; grep -E "^##|^large|^normal|max" gpurun_out/trip_a.log
