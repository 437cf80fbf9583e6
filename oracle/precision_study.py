"""ANALYSIS TOOL (test infrastructure, CPU only) - which operand formats can meet the 1e-3 contract, and at what tensor cost.

Runs the oracle network with every conv's operands rounded the way a tensor-core scheme would round them (fp32
accumulation, which is what the UMMA accumulators do) and prints max|out - fp32 reference|.  It backs the precision
section of DESIGN.md: the three-product bf16 split is the cheapest scheme that keeps a safe margin on the hard synthetic
case (recipe B at 512x512: activations up to 137, outputs saturating at +-1).

    python oracle/precision_study.py [scheme ...]        # default: all schemes on the three 256x256 cases
    python oracle/precision_study.py --hard [scheme ...] # add large/B at 512x512 (about a minute per scheme)

schemes (tensor cost in bf16-MMA equivalents per K step):
  bf16        1.0   bf16 x bf16                                   (FAST mode)
  f16         1.0   fp16 x fp16
  f16_a2      2.0   fp16 hi+lo activations x fp16 weights
  f16_f8corr  2.0   fp16 x fp16 + one K-doubled fp8 (e4m3) MMA for a_lo*w_hi + a_hi*w_lo, power-of-two scales
  f16_f8e5m2  2.0   same with e5m2 activations (range-safe, 3-bit significand)
  bf16x3      3.0   bf16 hi*hi + hi*lo + lo*hi                      (PARITY mode of round 1)
  f16x3       3.0   fp16 hi*hi + hi*lo + lo*hi                      (22-bit operands, same tensor cost: PARITY mode since round 2)

    python oracle/precision_study.py --mix                # round-2 question: can SOME layers run a 2-pass scheme?
runs large/B at 512x512 (the case with the least margin) with the 2-pass scheme f16_f8corr on one class of layers at a
time (and on growing unions of classes) while every other layer keeps bf16x3 or f16x3, and prints the error next to the
fraction of the tensor work that would become cheaper.
"""
from __future__ import annotations

import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import f2f_oracle as O  # noqa: E402

_conv2d = F.conv2d


def _q(t, dtype):
    return t.to(dtype).to(torch.float32)


def _q8(t, log2_scale, dtype):
    lim = 448.0 if dtype == torch.float8_e4m3fn else 57344.0
    s = 2.0 ** log2_scale
    return (t * s).clamp(-lim, lim).to(dtype).to(torch.float32) / s


def scheme_conv(scheme: str):
    def conv(x, w, b, stride, pad):
        c = lambda a, ww: _conv2d(a, ww, None, stride, pad)  # noqa: E731
        if scheme == "bf16":
            return c(_q(x, torch.bfloat16), _q(w, torch.bfloat16))
        if scheme == "f16":
            return c(_q(x, torch.float16), _q(w, torch.float16))
        if scheme == "bf16x3":
            xh, wh = _q(x, torch.bfloat16), _q(w, torch.bfloat16)
            return c(xh, wh) + c(xh, _q(w - wh, torch.bfloat16)) + c(_q(x - xh, torch.bfloat16), wh)
        if scheme == "f16x3":
            xh, wh = _q(x, torch.float16), _q(w, torch.float16)
            return c(xh, wh) + c(xh, _q(w - wh, torch.float16)) + c(_q(x - xh, torch.float16), wh)
        if scheme == "f16_a2":
            xh, wh = _q(x, torch.float16), _q(w, torch.float16)
            return c(xh, wh) + c(_q(x - xh, torch.float16), wh)
        if scheme in ("f16_f8corr", "f16_f8e5m2"):
            xh, wh = _q(x, torch.float16), _q(w, torch.float16)
            ka = torch.float8_e4m3fn if scheme == "f16_f8corr" else torch.float8_e5m2
            # scales chosen so that both correction products carry 2^22 and activations up to ~250 do not saturate e4m3
            return c(xh, wh) + c(_q8(x - xh, 14, ka), _q8(w, 8, torch.float8_e4m3fn)) + \
                c(_q8(x, 0, ka), _q8(w - wh, 22, torch.float8_e4m3fn))
        raise ValueError(scheme)
    return conv


def layer_class(x, w, stride) -> str:
    """Class of a conv by its shape: 'CinxCout@Hout' (the groups of SURVEY.md section 8a)."""
    return f"{w.shape[1]}->{w.shape[0]}@{x.shape[2] // stride}"


def mixed_conv(cheap: str, base: str, cheap_classes, seen: dict):
    cc, cb = scheme_conv(cheap), scheme_conv(base)

    def conv(x, w, b, stride, pad):
        k = layer_class(x, w, stride)
        macs = x.shape[0] * (x.shape[2] // stride) * (x.shape[3] // stride) * w.shape[0] * w.shape[1] * 9
        seen[k] = seen.get(k, 0) + macs
        return (cc if k in cheap_classes else cb)(x, w, b, stride, pad)
    return conv


def run_mixed(variant, recipe, size, cheap, base, cheap_classes):
    sd = O.make_state_dict(variant, recipe)
    fm, cand = O.make_inputs(1, size, size)
    x = torch.cat([fm, cand], 1)
    ref = O.generator_forward(sd, x, variant)
    seen = {}
    F.conv2d = mixed_conv(cheap, base, set(cheap_classes), seen)
    try:
        out = O.generator_forward(sd, x, variant)
    finally:
        F.conv2d = _conv2d
    tot = sum(seen.values())
    frac = sum(v for k, v in seen.items() if k in cheap_classes) / tot
    return (out - ref).abs().max().item(), frac, seen


def mix_study() -> None:
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    v, r, n = "large", "B", 512
    res = ["64->64@256", "128->128@128", "256->256@64", "512->512@32"]
    ups = ["1024->256@64", "512->128@128", "256->64@256", "1024->512@32"]
    for base in ("bf16x3", "f16x3"):
        e, _, seen = run_mixed(v, r, n, "f16_f8corr", base, [])
        print(f"base {base:7s} everywhere                         : max|err| = {e:.3e}", flush=True)
        for group in [[c] for c in res] + [ups, res[1:3], res[1:3] + ups, res, res + ups, list(seen.keys())]:
            e, frac, _ = run_mixed(v, r, n, "f16_f8corr", base, group)
            name = "+".join(group) if len(group) <= 3 else f"{len(group)} classes incl. {group[0]}"
            print(f"base {base:7s} + 2-pass on {name:44s}: max|err| = {e:.3e}   ({100 * frac:.0f} % of the MACs at 2 passes)", flush=True)


def run(variant: str, recipe: str, size: int, scheme: str) -> float:
    sd = O.make_state_dict(variant, recipe)
    fm, cand = O.make_inputs(1, size, size)
    x = torch.cat([fm, cand], 1)
    ref = O.generator_forward(sd, x, variant)
    F.conv2d = scheme_conv(scheme)          # the oracle calls torch.nn.functional.conv2d through the same module object
    try:
        out = O.generator_forward(sd, x, variant)
    finally:
        F.conv2d = _conv2d
    return (out - ref).abs().max().item()


def main() -> None:
    if "--mix" in sys.argv:
        mix_study()
        return
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    schemes = args or ["bf16", "f16", "f16_a2", "f16_f8corr", "f16_f8e5m2", "bf16x3"]
    cases = [("large", "A", 256), ("large", "B", 256), ("normal", "B", 256)]
    if "--hard" in sys.argv:
        cases.append(("large", "B", 512))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for s in schemes:
        for v, r, n in cases:
            t = time.time()
            print(f"{s:11s} {v:6s} {r} {n}: max|err| = {run(v, r, n, s):.3e}   [{time.time() - t:.1f} s]", flush=True)


if __name__ == "__main__":
    main()
