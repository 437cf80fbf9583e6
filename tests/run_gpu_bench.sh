#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 308 -c 154 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_list.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 304 -c 14 -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; tail -3 gpurun_out/ncu_list.log; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out
