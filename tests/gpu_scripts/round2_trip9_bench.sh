#!/bin/bash
# Round-2 closing trip: the bench line exactly as the driver runs it (default flags and the driver's --steps 20 --warmup 3 form)
# and the ncu launch list of that command.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_driver_form.json 2> gpurun_out/bench_driver_form.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 320 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/ncu_list.log 2>&1
cut -c1-400 gpurun_out/bench.json; tail -2 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_driver_form.json
