#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python tests/gpu_bringup.py layers large B parity 512 2 | grep -E "BAD|final|FAILED|bad fraction|channels|per-row|per-col" | head -20
timeout 600 python tests/gpu_bringup.py layers normal B fast 512 4 | grep -E "BAD|final|FAILED" | head -8
timeout 300 python tests/gpu_bringup.py time large A parity 512 16
LSPG_NO_RESIDENT=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 16
timeout 300 python tests/gpu_bringup.py time large A fast 512 16
timeout 300 python tests/gpu_bringup.py time large A parity 512 1
} > gpurun_out/bringup16.log 2>&1
tail -14 gpurun_out/bringup16.log
