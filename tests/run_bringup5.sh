#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tests/gpu_bringup.py final large B parity 512 2
timeout 300 python tests/gpu_bringup.py final normal A fast 256 3
timeout 300 python tests/gpu_bringup.py time large A fast 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 1
timeout 300 python tests/gpu_bringup.py time large A fast 512 1
LSPG_NO_GRAPH=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 8
LSPG_NO_GRAPH=1 timeout 300 python tests/gpu_bringup.py time large A fast 512 1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
} > gpurun_out/bringup5.log 2>&1
tail -c 2500 gpurun_out/bringup5.log
