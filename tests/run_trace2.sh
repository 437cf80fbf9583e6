#!/bin/bash
mkdir -p gpurun_out
{
for L in 1 2 70 75 0; do
LSPG_TRACE_LAYER=$L timeout 300 python tests/gpu_trace.py large parity 16 | grep -E "^layer|CTA 1:|   tile [1-4]:|mma:|epi:|wait_acc|total"
done
} > gpurun_out/trace2.log 2>&1
cat gpurun_out/trace2.log
