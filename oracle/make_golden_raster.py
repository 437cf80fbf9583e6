"""Generate tests/golden/raster_*.npz by running the UNMODIFIED drawing methods of the reference.  TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference and cv2):  ``python oracle/make_golden_raster.py``

``datasets/face_dataset.py`` cannot be imported here (albumentations, skimage ... are not installed), but the three methods on
the rasteriser path need only numpy and cv2.  They are lifted out of the reference file with ``ast`` at run time -
``FaceDataset.get_feature_image`` (:285-297), ``draw_shoulder_points`` (:300-309), ``draw_face_feature_maps`` (:312-323) and
the ``self.part_list`` assignment of ``__init__`` (:34-42) - compiled into a bare class and executed unchanged (nothing is
copied into this repository).  Stored per case: the float32 landmark / shoulder tracks and the bit-packed uint8 maps.
The restatement in oracle/raster_oracle.py is asserted equal while generating.
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import raster_oracle as R  # noqa: E402

REF_FILE = "/root/reference/datasets/face_dataset.py"

# (name, batch, (W, H), seed, spill, with shoulders)
CASES = [
    ("raster_face_512", 3, (512, 512), 11, 0.0, False),
    ("raster_face_shoulders_512", 3, (512, 512), 12, 0.0, True),
    ("raster_spill_256", 4, (256, 256), 13, 0.35, True),
    ("raster_wide_384x256", 2, (384, 256), 14, 0.1, True),
]


def reference_drawer():
    tree = ast.parse(open(REF_FILE).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "FaceDataset")
    wanted = ("get_feature_image", "draw_shoulder_points", "draw_face_feature_maps")
    funcs = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in wanted]
    assert len(funcs) == len(wanted)
    init = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "__init__")
    part = [s for s in init.body if isinstance(s, ast.Assign)
            and any(isinstance(t, ast.Attribute) and t.attr == "part_list" for t in s.targets)]
    assert len(part) == 1
    new_init = ast.parse("def __init__(self):\n    pass").body[0]
    new_init.body = part
    new_cls = ast.ClassDef(name="RefDrawer", bases=[], keywords=[], body=[new_init] + funcs, decorator_list=[])
    if "type_params" in ast.ClassDef._fields:
        new_cls.type_params = []
    mod = ast.Module(body=ast.parse("import cv2\nimport numpy as np").body + [new_cls], type_ignores=[])
    ns: dict = {}
    exec(compile(ast.fix_missing_locations(mod), REF_FILE, "exec"), ns)
    return ns["RefDrawer"]()


def main() -> None:
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    ref = reference_drawer()
    assert [list(map(list, e)) for e in ref.part_list] == [list(map(list, e)) for e in R.PART_LIST]
    for name, batch, size, seed, spill, with_sh in CASES:
        lm, sh = R.make_landmarks(batch, size, seed=seed, spill=spill)
        maps = []
        for b in range(batch):
            # the reference mutates its shoulder array only when image_pad is given; pass copies anyway
            img = ref.get_feature_image(lm[b].copy(), size, sh[b].copy() if with_sh else None, None)
            mine = R.draw_feature_map(lm[b], size, sh[b] if with_sh else None)
            assert img.dtype == np.uint8 and img.shape == (size[1], size[0])
            assert np.array_equal(img, mine), f"{name}[{b}]: restatement differs from the reference drawing"
            maps.append(np.packbits(img > 0))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), landmarks=lm, shoulders=sh if with_sh else np.zeros((0,), np.float32),
                            size=np.array(size), packed=np.stack(maps))
        print(f"{name}: {batch} maps {size}, coverage {np.mean([np.unpackbits(m).mean() for m in maps]):.4f}")


if __name__ == "__main__":
    main()
