#!/bin/bash
mkdir -p gpurun_out
export LSPG_PER_LAYER=1
{
timeout 300 python tests/gpu_bringup.py layers large B parity 512 2 | grep -v " ok $" | tail -5
timeout 300 python tests/gpu_bringup.py time large A fast 512 8
LSPG_NO_PATCH=1 timeout 300 python tests/gpu_bringup.py time large A fast 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 8
LSPG_NO_PATCH=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 8
} > gpurun_out/bringup3.log 2>&1
tail -c 3000 gpurun_out/bringup3.log
