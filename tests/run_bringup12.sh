#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tests/gpu_bringup.py layers large B parity 512 1 | grep -E "BAD|final|FAILED" | head -8
timeout 300 python tests/gpu_bringup.py layers normal B fast 256 3 | grep -E "BAD|final|FAILED" | head -8
timeout 300 python tests/gpu_bringup.py time large A parity 512 16
timeout 300 python tests/gpu_bringup.py time large A parity 512 1
timeout 300 python tests/gpu_bringup.py time large A fast 512 1
LSPG_SPLITK_TWO_PASS=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 1
timeout 300 python tests/gpu_bringup.py time large A parity 512 8
} > gpurun_out/bringup12.log 2>&1
cat gpurun_out/bringup12.log | tail -30
