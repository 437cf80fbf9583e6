"""Dump the in-kernel clock64 trace of one conv layer (debug tool; see conv_umma.cuh kTraceSlots).
    LSPG_TRACE_LAYER=<i> python tests/gpu_trace.py <variant> <mode> <B>"""
import ctypes as C
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import f2f_oracle as O
from livespeechportraits_b200 import _lib
from livespeechportraits_b200.generator import Feature2Face_G

variant, mode, B = sys.argv[1], sys.argv[2], int(sys.argv[3])
os.environ["LSPG_NO_GRAPH"] = "1"
net = Feature2Face_G(types.SimpleNamespace(isTrain=False, size=variant, n_downsample_G=8, ngf=64, fp16=0), precision=mode)
net.load_state_dict(O.make_state_dict(variant, "A"))
net = net.cuda().eval()
fm, cand = O.make_inputs(B, 512, 512)
x = torch.cat([fm, cand], 1).cuda()
for _ in range(3):
    net(x)
torch.cuda.synchronize()
SLOTS, TILES = 4 + 8 * 12, 12
buf = np.zeros(256 * SLOTS, np.uint64)
_lib.check(net._lib.lspg_debug_read_trace(net._handle, buf.ctypes.data, buf.size))
t = buf.reshape(256, SLOTS).astype(np.int64)
layer = int(os.environ["LSPG_TRACE_LAYER"])
row = net.layer_table(512, 512)[layer]
print(f"layer {layer}: {row}")
for cta in (0, 1, 73, 147):
    e = t[cta]
    if e[0] == 0:
        continue
    t0 = e[0]
    print(f"CTA {cta}: prologue {e[1] - t0} cyc")
    for i in range(TILES):
        s = e[4 + 8 * i: 12 + 8 * i]
        if s[0] == 0 and s[3] == 0:
            break
        rel = [int(v - t0) if v else -1 for v in s]
        print(f"   tile {i}: mma_start {rel[0]:7d} first_B {rel[1]:7d} mma_issued {rel[2]:7d} | epi_wait {rel[3]:7d} acc_ready {rel[4]:7d} "
              f"epi_done {rel[5]:7d} | prod_first {rel[6]:7d} prod_last {rel[7]:7d}")
# aggregate: average per-tile MMA issue span, epilogue span, acc_ready->epi_done
import statistics
spans = {"mma": [], "epi": [], "wait_acc": [], "total": []}
for cta in range(256):
    e = t[cta]
    if e[0] == 0:
        continue
    last = 0
    for i in range(TILES):
        s = e[4 + 8 * i: 12 + 8 * i]
        if s[0] == 0:
            break
        spans["mma"].append(int(s[2] - s[0]))
        spans["epi"].append(int(s[5] - s[4]))
        spans["wait_acc"].append(int(s[4] - s[3]))
        last = max(last, int(s[5] - e[0]))
    spans["total"].append(last)
for k, v in spans.items():
    if v:
        print(f"{k}: n={len(v)} mean={statistics.mean(v):.0f} median={statistics.median(v):.0f} max={max(v)}")
