#!/bin/bash
# 2-GPU sanity run of bench.py (same launch line as the driver uses) + the reference arm under torchrun.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 40 --warmup 3 --no-extras > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_n1_short.json 2> gpurun_out/bench_n1_short.err
tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json; cat gpurun_out/bench_n1_short.json | cut -c1-400
