#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?" >> gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err; cat gpurun_out/bench_ref_n2.json
