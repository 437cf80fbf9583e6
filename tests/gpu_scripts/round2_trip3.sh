#!/bin/bash
# Round-2 trip 3 (1 GPU): fp16 PARITY limbs + cluster split-K + fault path: parity suite, bench, batch-1 A/B of the cluster split.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
for B in 1 32; do
  for i in 1 2; do
    echo "## cluster-split B$B" >> gpurun_out/ab_csplit.log; LSPG_CLUSTER_SPLIT=1 timeout 200 python tests/gpu_bringup.py time large A parity 512 $B >> gpurun_out/ab_csplit.log 2>&1
    echo "## finisher B$B" >> gpurun_out/ab_csplit.log; LSPG_CLUSTER_SPLIT=0 timeout 200 python tests/gpu_bringup.py time large A parity 512 $B >> gpurun_out/ab_csplit.log 2>&1
  done
done
LSPG_PER_LAYER=1 timeout 300 python tests/gpu_bringup.py time large A parity 512 1 > gpurun_out/b1_layers.log 2>&1
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 160 --csv --log-file gpurun_out/launches_b1.csv \
    python tests/gpu_bringup.py time large A parity 512 1 > gpurun_out/ncu_list_b1.log 2>&1
tail -30 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; grep -E "##|ms/forward" gpurun_out/ab_csplit.log; cut -c1-500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
