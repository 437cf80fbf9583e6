"""CPU emulation of the *launch plan* the C library builds (test infrastructure, not a product path).

Executes, with torch on the CPU, exactly what the CUDA kernels are told to do: the layer list, tap
tables (parity views for stride 2, output phases for the folded upsample, concat as K ranges), the
packed K-major weights and the folded scale/shift, all read back through the C ABI's introspection
calls of a host-only handle.  Comparing its output with the oracle validates the host logic (network
structure, key mapping, weight packing, BatchNorm folding, tap geometry) without a GPU; what remains
for the GPU tests is the PTX machinery itself.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch

from livespeechportraits_b200 import _lib

KIND_HEAD, KIND_S1, KIND_S2, KIND_UP, KIND_TAIL = range(5)


def bf16_bits_to_f32(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy((a.astype(np.uint32) << 16).view(np.float32).copy())


class HostPlan:
    def __init__(self, variant: str, state_dict: Dict[str, torch.Tensor], ngf: int = 64, num_downs: int = 8):
        self.lib = _lib.load()
        self.h = C.c_void_p()
        _lib.check(self.lib.lspg_create(C.byref(self.h), _lib.LSPG_VARIANT[variant], ngf, num_downs, 13, 3, -1))
        keep, names = [], []
        for k, v in state_dict.items():
            if k.endswith("num_batches_tracked"):
                continue
            names.append(k.encode())
            keep.append(v.detach().float().contiguous())
        arr = (_lib.LspgTensor * len(keep))()
        for i, (nm, t) in enumerate(zip(names, keep)):
            arr[i].name, arr[i].numel = nm, t.numel()
            arr[i].data = C.cast(t.data_ptr(), C.POINTER(C.c_float))
        _lib.check(self.lib.lspg_load_weights(self.h, arr, len(keep)))
        n = C.c_int()
        _lib.check(self.lib.lspg_num_layers(self.h, C.byref(n)))
        self.layers: List[_lib.LspgLayerInfo] = []
        for i in range(n.value):
            info = _lib.LspgLayerInfo()
            _lib.check(self.lib.lspg_layer_info_get(self.h, i, C.byref(info)))
            self.layers.append(info)

    def close(self):
        if self.h:
            self.lib.lspg_destroy(self.h)
            self.h = C.c_void_p()

    PARITY_WEIGHT_SCALE = 256.0          # include/lspg.h: LSPG_PARITY_WEIGHT_SCALE

    def packed(self, i: int, limbs: int) -> torch.Tensor:
        """Weights as the kernels multiply them: ``limbs == 2`` = PARITY (fp16 hi + lo limbs of w * 256, scaled back here),
        ``limbs == 1`` = FAST (bf16)."""
        L = self.layers[i]
        count = L.n_phases * L.cout_pad * L.k_total
        buf = np.zeros(count, np.uint16)
        if limbs == 1:
            _lib.check(self.lib.lspg_layer_packed(self.h, i, 2, buf.ctypes.data, count))
            out = bf16_bits_to_f32(buf)
        else:
            out = torch.zeros(count)
            for limb in range(2):
                _lib.check(self.lib.lspg_layer_packed(self.h, i, limb, buf.ctypes.data, count))
                out += torch.from_numpy(buf.view(np.float16).astype(np.float32))
            out = out / self.PARITY_WEIGHT_SCALE
        return out.view(L.n_phases, L.cout_pad, L.k_total)

    def affine(self, i: int):
        L = self.layers[i]
        s = np.zeros(L.cout_pad, np.float32)
        b = np.zeros(L.cout_pad, np.float32)
        _lib.check(self.lib.lspg_layer_affine(self.h, i, s.ctypes.data, b.ctypes.data, L.cout_pad))
        return torch.from_numpy(s), torch.from_numpy(b)


def _shift(a: torch.Tensor, dy: int, dx: int) -> torch.Tensor:
    """b[n, y, x] = a[n, y+dy, x+dx], zero outside (what TMA's out-of-bounds fill gives the kernel)."""
    n, h, w, c = a.shape
    out = torch.zeros_like(a)
    ys0, ys1 = max(0, dy), min(h, h + dy)
    xs0, xs1 = max(0, dx), min(w, w + dx)
    if ys1 > ys0 and xs1 > xs0:
        out[:, ys0 - dy:ys1 - dy, xs0 - dx:xs1 - dx] = a[:, ys0:ys1, xs0:xs1]
    return out


def pack_input_s2d(x: torch.Tensor) -> torch.Tensor:
    """aux_kernels.cuh:pack_input_s2d_kernel in torch: [B,13,H,W] -> [B,H/2,W/2,64]."""
    b, c, h, w = x.shape
    s = torch.zeros(b, h // 2, w // 2, 64)
    for py in range(2):
        for px in range(2):
            s[..., (py * 2 + px) * 16:(py * 2 + px) * 16 + c] = x[:, :, py::2, px::2].permute(0, 2, 3, 1)
    return s


def run_plan(plan: HostPlan, x: torch.Tensor, limbs: int = 2, round_act=None,
             taps_out: Optional[Dict[int, torch.Tensor]] = None) -> torch.Tensor:
    """Execute the plan on [B,13,H,W] fp32 input.  ``limbs``: 1 = FAST (bf16 weights), 2 = PARITY (fp16 hi+lo weights).
    ``round_act``: optional function applied to every stored activation (e.g. bf16 rounding)."""
    q = round_act or (lambda t: t)
    tensors: Dict[int, torch.Tensor] = {0: q(pack_input_s2d(x))}
    result = None
    for i, L in enumerate(plan.layers):
        W = plan.packed(i, limbs)                                 # [phases, cout_pad, K]
        scale, shift = plan.affine(i)
        srcs = [tensors[L.src[s]] for s in range(L.n_src)]
        kb_per_tap = sum(int(L.cin[s]) for s in range(L.n_src))
        b = srcs[0].shape[0]
        if L.kind == KIND_S2:
            hs, ws = srcs[0].shape[1] // 2, srcs[0].shape[2] // 2
        else:
            hs, ws = srcs[0].shape[1], srcs[0].shape[2]
        phase_out = []
        for z in range(L.n_phases):
            acc = torch.zeros(b, hs, ws, L.cout_pad)
            for t in range(L.n_taps):
                dx, dy, amap = int(L.tap_dx[z][t]), int(L.tap_dy[z][t]), int(L.tap_map[z][t])
                koff = t * kb_per_tap
                for s in range(L.n_src):
                    a = srcs[s]
                    if L.kind == KIND_S2:
                        a = a[:, (amap >> 1)::2, (amap & 1)::2]    # parity view
                    a = _shift(a, dy, dx)
                    cin = int(L.cin[s])
                    acc += torch.einsum("nyxc,oc->nyxo", a, W[z, :, koff:koff + cin])
                    koff += cin
            phase_out.append(acc * scale + shift)
        if L.kind == KIND_TAIL:
            y = torch.tanh(phase_out[0])                           # [B,hs,ws,16]; col = (py*2+px)*3 + c
            out = torch.zeros(b, 3, 2 * hs, 2 * ws)
            for py in range(2):
                for px in range(2):
                    for c in range(3):
                        out[:, c, py::2, px::2] = y[..., (py * 2 + px) * 3 + c]
            result = out
            continue
        if L.kind == KIND_UP:
            y = torch.zeros(b, 2 * hs, 2 * ws, L.cout_pad)
            for z in range(4):
                y[:, (z >> 1)::2, (z & 1)::2] = phase_out[z]
        else:
            y = phase_out[0]
        if L.res >= 0:
            y = y + tensors[L.res]
        if L.relu:
            y = torch.relu(y)
        tensors[L.out] = q(y)
        if taps_out is not None:
            taps_out[L.out] = tensors[L.out]
    return result
