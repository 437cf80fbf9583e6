"""Generate tests/golden/*.npz by running the UNMODIFIED reference module.  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference):  ``python oracle/make_golden.py``

For every case the synthetic weights of ``f2f_oracle.make_state_dict`` are loaded into the real
``models.feature2face_G.Feature2Face_G`` with ``strict=True`` and the module is run on the synthetic
inputs of ``f2f_oracle.make_inputs`` through ``torch.cat([fm, cand], 1)`` exactly as
``Feature2FaceModel.inference`` does (models/feature2face_model.py:225-237).  Stored per case: a
strided sub-sample of the fp32 output, its full-tensor statistics (sum, sum of squares, min, max as
float64) and sub-samples of two intermediate activations, so the fixtures stay small.
The GPU box has no reference checkout: tests there regenerate weights/inputs from the same seeds and
compare against these files.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import f2f_oracle as O  # noqa: E402

# (name, variant, recipe, batch, H, W, output stride)
CASES = [
    ("large_A_b1_256", "large", "A", 1, 256, 256, 2),
    ("large_B_b1_256", "large", "B", 1, 256, 256, 2),
    ("normal_A_b1_256", "normal", "A", 1, 256, 256, 2),
    ("normal_B_b2_256", "normal", "B", 2, 256, 256, 2),
    ("large_A_b1_512", "large", "A", 1, 512, 512, 8),
    ("normal_B_b1_512", "normal", "B", 1, 512, 512, 8),
    ("large_B_b1_512", "large", "B", 1, 512, 512, 8),
]


def stats(t: torch.Tensor) -> np.ndarray:
    d = t.double()
    return np.array([d.sum().item(), (d * d).sum().item(), d.min().item(), d.max().item()], np.float64)


def main() -> None:
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    nets = {}
    for name, variant, recipe, batch, h, w, stride in CASES:
        if variant not in nets:
            nets[variant] = O.reference_generator(variant)
        net = nets[variant]
        sd = O.make_state_dict(variant, recipe)
        net.load_state_dict(sd, strict=True)
        fm, cand = O.make_inputs(batch, h, w)
        with torch.no_grad():
            ref = net(torch.cat([fm, cand], 1))
        taps = {}
        mine = O.generator_forward(sd, torch.cat([fm, cand], 1), variant, taps=taps)
        err = (ref - mine).abs().max().item()
        assert err <= 2e-6, f"{name}: restatement differs from the reference by {err}"
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            variant=variant, recipe=recipe, batch=batch, height=h, width=w, stride=stride,
            out_sub=ref[:, :, ::stride, ::stride].numpy(), out_stats=stats(ref),
            e1_sub=taps["e1"][:, ::8, ::stride * 2, ::stride * 2].numpy(), e1_stats=stats(taps["e1"]),
            d1_sub=taps["d1"][:, ::8, ::stride * 2, ::stride * 2].numpy(), d1_stats=stats(taps["d1"]),
            torch_version=torch.__version__,
        )
        print(f"{name}: out std {ref.std().item():.4f} max|ref-oracle| {err:.2e}")
    # Known-answer facts of the reference: zero in -> zero out under recipe A.
    net = nets["normal"]
    net.load_state_dict(O.make_state_dict("normal", "A"), strict=True)
    with torch.no_grad():
        z = net(torch.zeros(1, 13, 256, 256))
    assert float(z.abs().max()) == 0.0
    print("zero-in/zero-out KAT holds on the reference")


if __name__ == "__main__":
    main()
