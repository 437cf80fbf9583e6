#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python tests/gpu_bringup.py layers large B parity 512 8 | grep -E "BAD|final|FAILED" | head -8
timeout 300 python tests/gpu_bringup.py time large A parity 512 16
timeout 300 python tests/gpu_bringup.py time large A parity 512 32
timeout 300 python tests/gpu_bringup.py time large A fast 512 32
timeout 300 python tests/gpu_bringup.py time normal A parity 512 32
timeout 300 python tests/gpu_bringup.py time large A parity 512 1
} > gpurun_out/bringup15.log 2>&1
tail -14 gpurun_out/bringup15.log
