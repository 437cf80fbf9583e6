#!/bin/bash
mkdir -p gpurun_out
{
LSPG_PER_LAYER=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 16
LSPG_PER_LAYER=1 timeout 300 python tests/gpu_bringup.py time large A fast 512 16
} > gpurun_out/bringup14.log 2>&1
tail -5 gpurun_out/bringup14.log
