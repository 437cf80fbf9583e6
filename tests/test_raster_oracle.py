"""CPU tests of the rasteriser oracle (N2): pinned against cv2 - the reference's own dependency for this step - and
against golden maps produced by the reference's unmodified drawing methods (oracle/make_golden_raster.py)."""
import glob
import os
import random

import numpy as np
import pytest

from oracle import raster_oracle as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
cv2 = pytest.importorskip("cv2")


def test_segment_tables_follow_the_reference_part_list():
    segs = R.face_segments()
    assert len(segs) == 72 and segs[0] == (0, 1) and segs[17] == (18, 18) and segs[-1] == (58, 46)
    assert R.shoulder_segments(18) == [(j, j + 1) for j in range(8)] + [(9 + j, 10 + j) for j in range(8)]


def test_thick_line_is_bit_exact_with_cv2_on_random_segments():
    rng = random.Random(5)
    for it in range(4000):
        w, h = rng.choice([(64, 64), (48, 80), (96, 40)])
        m = rng.choice([0, 0, 5, 40])                                  # how far endpoints may leave the image
        p0 = (rng.randint(-m, w - 1 + m), rng.randint(-m, h - 1 + m))
        if it % 5 == 0:                                                # degenerate and very short segments
            p1 = (p0[0] + rng.randint(-3, 3), p0[1] + rng.randint(-3, 3))
        else:
            p1 = (rng.randint(-m, w - 1 + m), rng.randint(-m, h - 1 + m))
        a = np.zeros((h, w), np.uint8)
        b = np.zeros((h, w), np.uint8)
        cv2.line(a, p0, p1, 255, 2)
        R.thick_line(b, p0, p1)
        assert np.array_equal(a, b), (w, h, p0, p1)


def test_edge_cases():
    for p0, p1 in [((10, 10), (10, 10)), ((0, 0), (63, 63)), ((-5, -5), (-1, -1)), ((-3, 10), (-3, 50)), ((63, 0), (63, 63)),
                   ((-100, 30), (200, 31)), ((30, -100), (31, 200)), ((-2, -2), (65, 65)), ((66, 10), (66, 20))]:
        a = np.zeros((64, 64), np.uint8)
        b = np.zeros((64, 64), np.uint8)
        cv2.line(a, p0, p1, 255, 2)
        R.thick_line(b, p0, p1)
        assert np.array_equal(a, b), (p0, p1)


def test_feature_map_matches_cv2_pipeline():
    for spill in (0.0, 0.3):
        lm, sh = R.make_landmarks(2, (256, 256), seed=21, spill=spill)
        for b in range(2):
            assert np.array_equal(R.draw_feature_map(lm[b], (256, 256), sh[b]), R.draw_feature_map_cv2(lm[b], (256, 256), sh[b]))
    fm = R.feature_map_tensor(lm[0], (256, 256), sh[0])
    assert fm.shape == (1, 256, 256) and fm.dtype == np.float32 and set(np.unique(fm)) <= {0.0, 1.0}


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "raster_*.npz"))), ids=os.path.basename)
def test_golden_maps_of_the_reference_drawing_code(path):
    g = np.load(path)
    w, h = (int(v) for v in g["size"])
    lm, sh = g["landmarks"], g["shoulders"]
    for b in range(lm.shape[0]):
        mine = R.draw_feature_map(lm[b], (w, h), sh[b] if sh.size else None)
        ref = np.unpackbits(g["packed"][b])[: w * h].reshape(h, w) * 255
        assert np.array_equal(mine, ref), (os.path.basename(path), b)


def test_goldens_present():
    assert len(glob.glob(os.path.join(GOLDEN, "raster_*.npz"))) == 4
