#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tests/gpu_bringup.py time large A fast 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 8
timeout 300 python tests/gpu_bringup.py time large A parity 512 1
timeout 300 python tests/gpu_bringup.py time large A fast 512 1
LSPG_NO_PDL=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 8
LSPG_NO_PDL=1 timeout 300 python tests/gpu_bringup.py time large A fast 512 1
} > gpurun_out/bringup6.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 260 --csv --log-file gpurun_out/launches2.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_list2.log 2>&1
tail -c 2500 gpurun_out/bringup6.log; tail -3 gpurun_out/ncu_list2.log
