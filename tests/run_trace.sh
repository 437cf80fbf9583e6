#!/bin/bash
mkdir -p gpurun_out
{
LSPG_TRACE_LAYER=6 timeout 300 python tests/gpu_trace.py large fast 8
LSPG_TRACE_LAYER=6 timeout 300 python tests/gpu_trace.py large parity 8
LSPG_TRACE_LAYER=1 timeout 300 python tests/gpu_trace.py large fast 8
LSPG_TRACE_LAYER=16 timeout 300 python tests/gpu_trace.py large fast 8
} > gpurun_out/trace.log 2>&1
cat gpurun_out/trace.log | tail -120
