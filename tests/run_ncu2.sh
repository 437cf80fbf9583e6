#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/gpu_bringup.py layers large B parity 512 1 | grep -E "BAD|final" | head -5 > gpurun_out/ncu2_check.log 2>&1
LSPG_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_patch -s 58 -c 10 -o gpurun_out/prof_fast \
    python tests/gpu_bringup.py time large A fast 512 8 > gpurun_out/ncu_fast.log 2>&1
LSPG_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_patch -s 58 -c 10 -o gpurun_out/prof_parity \
    python tests/gpu_bringup.py time large A parity 512 8 > gpurun_out/ncu_parity.log 2>&1
cat gpurun_out/ncu2_check.log; tail -3 gpurun_out/ncu_fast.log; ls -la gpurun_out/*.ncu-rep
