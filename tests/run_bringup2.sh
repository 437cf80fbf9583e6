#!/bin/bash
mkdir -p gpurun_out
{
echo "##### default (base offset on)"
timeout 300 python tests/gpu_bringup.py layers normal B fast 256 1 | grep -v " ok $" | tail -40
echo "--- exit $?"
echo "##### LSPG_NO_BASE_OFFSET=1"
LSPG_NO_BASE_OFFSET=1 timeout 300 python tests/gpu_bringup.py layers normal B fast 256 1 | grep -v " ok $" | tail -40
echo "--- exit $?"
echo "##### parity default"
timeout 300 python tests/gpu_bringup.py layers large B parity 512 2 | grep -v " ok $" | tail -40
echo "--- exit $?"
timeout 300 python tests/gpu_bringup.py time large A fast 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 8
LSPG_NO_PATCH=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 1
timeout 300 python tests/gpu_bringup.py time large A fast 512 1
} > gpurun_out/bringup2.log 2>&1
tail -c 7000 gpurun_out/bringup2.log
