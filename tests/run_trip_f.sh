#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
{
T="timeout 300 python tests/gpu_bringup.py"
echo "## final B32";               $T final large A parity 512 32
echo "## B32";                     $T time large A parity 512 32
echo "## B32 nowave"; LSPG_NO_WAVE_RULE=1 $T time large A parity 512 32
echo "## B37";                     $T time large A parity 512 37
echo "## B32";                     $T time large A parity 512 32
echo "## B32 nowave"; LSPG_NO_WAVE_RULE=1 $T time large A parity 512 32
echo "## B37";                     $T time large A parity 512 37
echo "## B16";                     $T time large A parity 512 16
echo "## B16 nowave"; LSPG_NO_WAVE_RULE=1 $T time large A parity 512 16
echo "## B74";                     $T time large A parity 512 74
echo "## B37 fast";                $T time large A fast 512 37
} > gpurun_out/trip_f.log 2>&1
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -4 gpurun_out/pytest_gpu.log; grep -E "^##|^large|^normal|max|pack" gpurun_out/trip_f.log | head -40; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
