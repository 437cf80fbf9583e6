#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tests/gpu_bringup.py layers large B parity 512 2 | grep -E "BAD|final|FAILED" | head -8
timeout 300 python tests/gpu_bringup.py layers normal B fast 256 3 | grep -E "BAD|final|FAILED" | head -8
timeout 300 python tests/gpu_bringup.py time large A fast 512 16
timeout 300 python tests/gpu_bringup.py time large A parity 512 16
timeout 300 python tests/gpu_bringup.py time large A parity 512 1
for L in 1 2; do
LSPG_TRACE_LAYER=$L timeout 300 python tests/gpu_trace.py large parity 16 | grep -E "^layer|mma:|epi:|wait_acc|total"
done
} > gpurun_out/bringup11.log 2>&1
cat gpurun_out/bringup11.log | tail -40
