#!/bin/bash
mkdir -p gpurun_out
{
for i in 1 2; do
LSPG_EPI_WARPS=4 timeout 300 python tests/gpu_bringup.py time large A parity 512 16
LSPG_EPI_WARPS=8 timeout 300 python tests/gpu_bringup.py time large A parity 512 16
done
LSPG_EPI_WARPS=4 timeout 300 python tests/gpu_bringup.py time large A fast 512 16
LSPG_EPI_WARPS=8 timeout 300 python tests/gpu_bringup.py time large A fast 512 16
LSPG_EPI_WARPS=4 LSPG_PER_LAYER=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 16
LSPG_EPI_WARPS=8 LSPG_PER_LAYER=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 16
} > gpurun_out/ab.log 2>&1
grep -E "^large" gpurun_out/ab.log
