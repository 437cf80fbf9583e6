#!/bin/bash
# GPU bring-up driver: each stage in its own process (a device trap poisons the context) under a timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
{
timeout 300 python tests/gpu_bringup.py layers normal B fast 256 1
echo "--- exit $?"
timeout 300 python tests/gpu_bringup.py layers normal B parity 256 1
echo "--- exit $?"
timeout 300 python tests/gpu_bringup.py final large A parity 512 1
echo "--- exit $?"
timeout 300 python tests/gpu_bringup.py final large A fast 512 2
echo "--- exit $?"
timeout 300 python tests/gpu_bringup.py time large A fast 512 1
timeout 300 python tests/gpu_bringup.py time large A fast 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 8
timeout 300 python tests/gpu_bringup.py time normal A fast 512 8
} > gpurun_out/bringup.log 2>&1
tail -c 6000 gpurun_out/bringup.log
