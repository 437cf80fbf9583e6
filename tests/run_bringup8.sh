#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tests/gpu_bringup.py layers large B parity 512 1 | grep -v " ok $" | tail -8
timeout 300 python tests/gpu_bringup.py layers normal B fast 256 3 | grep -v " ok $" | tail -8
timeout 300 python tests/gpu_bringup.py time large A fast 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 32
timeout 300 python tests/gpu_bringup.py time large A fast 512 32
timeout 300 python tests/gpu_bringup.py time large A parity 512 1
timeout 300 python tests/gpu_bringup.py time large A fast 512 1
LSPG_PER_LAYER=1 timeout 300 python tests/gpu_bringup.py time large A fast 512 8
LSPG_PER_LAYER=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 8
} > gpurun_out/bringup8.log 2>&1
grep -E "^(large|normal|final|layer|===|FORWARD)" gpurun_out/bringup8.log | head -40
