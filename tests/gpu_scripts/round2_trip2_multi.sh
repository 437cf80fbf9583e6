#!/bin/bash
# Round-2 multi-GPU trip (gpurun --gpus N): ShardedRenderer on NCCL / symmetric memory, then bench.py as the driver launches it.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
LSPG_PER_LAYER=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 1 > gpurun_out/b1_layers.log 2>&1
NCCL_DEBUG=WARN timeout 600 $TR --master-port 29533 tests/gpu_scripts/sharded_check.py > gpurun_out/sharded_check_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/sharded_check_n$N.log
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --clip-frames 10000 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --gather nccl --no-extras > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err; echo "rc=$?" >> gpurun_out/bench_n${N}_nccl.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_n1_short.json 2> gpurun_out/bench_n1_short.err
grep -E "SHARDED_CHECK|FAILED|rank 0\]|rc=" gpurun_out/sharded_check_n$N.log | tail -20
for f in gpurun_out/bench_n$N.json gpurun_out/bench_n${N}_nccl.json gpurun_out/bench_n1_short.json; do echo "== $f"; cut -c1-700 $f; done
tail -3 gpurun_out/bench_n$N.err; tail -25 gpurun_out/pytest_gpu.log; head -3 gpurun_out/b1_layers.log
