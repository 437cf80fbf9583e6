#!/bin/bash
# Round-2 multi-GPU trip (gpurun --gpus N): ShardedRenderer in every gather mode on hardware, then bench.py as the driver launches it.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
NCCL_DEBUG=WARN timeout 240 $TR --master-port 29533 tests/gpu_scripts/sharded_check.py > gpurun_out/sharded_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/sharded_check_n$N.log
grep -E "SHARDED_CHECK|FAILED|rank 0\]|rc=" gpurun_out/sharded_check_n$N.log | tail -24
timeout 420 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --clip-frames 10000 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --gather nccl --no-extras > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "rc=$?" >> gpurun_out/bench_n${N}_nccl.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_n1_short.json 2> gpurun_out/bench_n1_short.err
for f in gpurun_out/bench_n$N.json gpurun_out/bench_n${N}_nccl.json gpurun_out/bench_n1_short.json; do echo "== $f"; cut -c1-400 $f; done
tail -4 gpurun_out/bench_n$N.err | cut -c1-300
