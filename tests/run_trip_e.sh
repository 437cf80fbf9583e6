#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
{
T="timeout 300 python tests/gpu_bringup.py"
echo "## B16";                     $T time large A parity 512 16
echo "## B32";                     $T time large A parity 512 32
echo "## B64";                     $T time large A parity 512 64
echo "## B16";                     $T time large A parity 512 16
echo "## B32";                     $T time large A parity 512 32
echo "## B24";                     $T time large A parity 512 24
echo "## B32 fast";                $T time large A fast 512 32
echo "## B32 normal";              $T time normal A parity 512 32
echo "## B32 per-layer"; LSPG_PER_LAYER=1 $T time large A parity 512 32
} > gpurun_out/trip_e.log 2>&1
timeout 600 python bench.py --batch 32 --steps 50 > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench rc=$?" >> gpurun_out/bench_b32.err
tail -4 gpurun_out/pytest_gpu.log; grep -E "^##|^large|^normal|max|pack" gpurun_out/trip_e.log | head -40; tail -3 gpurun_out/bench_b32.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b32.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
