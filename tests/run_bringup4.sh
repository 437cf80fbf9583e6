#!/bin/bash
mkdir -p gpurun_out
export LSPG_PER_LAYER=1
{
echo "##### split-K correctness (B=1 and B=8 trigger different splits)"
timeout 300 python tests/gpu_bringup.py layers large B parity 512 1 | grep -v " ok $" | tail -12
timeout 300 python tests/gpu_bringup.py layers normal B fast 256 3 | grep -v " ok $" | tail -12
echo "##### timing with split-K"
timeout 300 python tests/gpu_bringup.py time large A fast 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 8
echo "##### tap rotate"
LSPG_TAP_ROTATE=1 timeout 300 python tests/gpu_bringup.py time large A fast 512 8
LSPG_TAP_ROTATE=1 timeout 300 python tests/gpu_bringup.py final large A parity 512 2
echo "##### B=1"
timeout 300 python tests/gpu_bringup.py time large A parity 512 1
timeout 300 python tests/gpu_bringup.py time large A fast 512 1
LSPG_NO_SPLITK=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 1 | head -3
} > gpurun_out/bringup4.log 2>&1
tail -c 1500 gpurun_out/bringup4.log
