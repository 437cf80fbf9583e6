#!/bin/bash
# Steady-state clock64 timelines of the 64-channel layers (pair kernel) + one full ncu capture of the pair kernels.
mkdir -p gpurun_out
{
for L in 1 2 70 6 12; do
LSPG_TRACE_SKIP=20 LSPG_TRACE_LAYER=$L timeout 300 python tests/gpu_trace.py large parity 16 | grep -E "^layer|CTA [01]:|   tile [0-9]+:|mma:|epi:|wait_acc|total"
done
LSPG_TRACE_SKIP=20 LSPG_TRACE_LAYER=1 timeout 300 python tests/gpu_trace.py large fast 16 | grep -E "^layer|CTA [01]:|   tile [0-9]+:|mma:|epi:|wait_acc|total"
} > gpurun_out/trace3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 90 -c 45 -o gpurun_out/prof_pair -f \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_pair.log 2>&1
tail -3 gpurun_out/ncu_pair.log | cut -c1-300
cat gpurun_out/trace3.log | head -150
