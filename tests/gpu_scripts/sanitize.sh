#!/bin/bash
# compute-sanitizer over one forward per kernel family (race / sync / memory checks of the hand-rolled
# mbarrier + TMEM pipelines).  `normal`, 256x256, batch 2 reaches conv_patch (head, tail), conv_pair<64 stacked>, conv_pair<128>,
# conv_umma + splitk_reduce and the input packer; plain stream launches (LSPG_NO_GRAPH) so that every launch is attributed.
# Usage under gpurun: bash tests/gpu_scripts/sanitize.sh [per-tool timeout seconds]
mkdir -p gpurun_out
T=${1:-300}
CS=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
for tool in memcheck synccheck racecheck; do
  echo "== $tool" > gpurun_out/sanitize_$tool.log
  LSPG_NO_GRAPH=1 timeout $T $CS --tool $tool --print-limit 20 --launch-timeout 120 \
      python tests/gpu_bringup.py final normal B parity 256 2 >> gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|max\|out-oracle\||rc=" gpurun_out/sanitize_$tool.log | tail -4
done
