#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python tests/gpu_bringup.py layers large B parity 512 4 | grep -E "BAD|final|FAILED|bad fraction|channels|per-row|per-col|sample" | head -30
timeout 600 python tests/gpu_bringup.py layers normal B fast 512 8 | grep -E "BAD|final|FAILED" | head -8
timeout 600 python tests/gpu_bringup.py final large A parity 512 16
timeout 300 python tests/gpu_bringup.py time large A parity 512 16
LSPG_NO_PAIR=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 16
timeout 300 python tests/gpu_bringup.py time large A fast 512 16
LSPG_NO_PAIR=1 timeout 300 python tests/gpu_bringup.py time large A fast 512 16
} > gpurun_out/bringup13.log 2>&1
cat gpurun_out/bringup13.log | tail -50
