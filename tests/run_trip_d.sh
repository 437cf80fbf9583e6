#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
{
T="timeout 300 python tests/gpu_bringup.py"
echo "## final";                    $T final large A parity 512 16
echo "## base";                     $T time large A parity 512 16
echo "## base";                     $T time large A parity 512 16
echo "## fast";                     $T time large A fast 512 16
echo "## B1";                       $T time large A parity 512 1
echo "## per-layer"; LSPG_PER_LAYER=1 $T time large A parity 512 16
} > gpurun_out/trip_d.log 2>&1
LSPG_TRACE_SKIP=20 LSPG_TRACE_LAYERS=2,1,0,75 LSPG_TRACE_CTAS=0,1 timeout 300 python tests/gpu_trace.py large parity 16 > gpurun_out/trace7.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -4 gpurun_out/pytest_gpu.log; grep -E "^##|^large|^normal|max" gpurun_out/trip_d.log | head -40; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], json.dumps(d.get('e2e_from_landmarks_uint8')), json.dumps(d.get('e2e_uint8_images')))"
