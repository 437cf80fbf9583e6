#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tests/gpu_bringup.py layers large B parity 512 2 | grep -E "BAD|final|FAILED" | head -8
timeout 300 python tests/gpu_bringup.py layers normal B fast 256 3 | grep -E "BAD|final|FAILED" | head -8
timeout 300 python tests/gpu_bringup.py time large A fast 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 8
LSPG_NO_CLUSTER=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 32
timeout 300 python tests/gpu_bringup.py time large A fast 512 32
LSPG_TRACE_LAYER=6 timeout 300 python tests/gpu_trace.py large parity 8 | tail -12
LSPG_TRACE_LAYER=6 timeout 300 python tests/gpu_trace.py large fast 8 | tail -12
} > gpurun_out/bringup10.log 2>&1
cat gpurun_out/bringup10.log | tail -60
