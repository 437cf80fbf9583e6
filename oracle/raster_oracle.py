"""TEST INFRASTRUCTURE - CPU restatement of the feature-map rasteriser that feeds the generator.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

What the reference does (datasets/face_dataset.py:276-323, called per frame from demo.py:262-265):
  * ``draw_face_feature_maps`` (:312-323): a zero uint8 H x W image; for every consecutive landmark pair of every edge list
    in ``part_list`` (:34-42) ``cv2.line(img, int(pt1), int(pt2), 255, 2)`` (Python ``int()``: truncation toward zero);
  * ``draw_shoulder_points`` (:300-309): two polylines over the two halves of the shoulder points, same call;
  * ``get_data_test_mode`` (:276-282): ``img[None].astype(float32) / 255`` -> a {0,1} map of shape [1,H,W].
All lines have the same colour, so the map is the union of the pixel sets of the individual ``cv2.line`` calls.

The arithmetic lives in a third-party dependency, OpenCV (``cv2.line`` with thickness 2, LINE_8, shift 0).  The reference pins
opencv_python==4.4.0.40 (requirements.txt) / 4.1.2.30 (cog.yaml); neither source tree is vendored.  This file restates the
published algorithm of imgproc/drawing.cpp (line -> ThickLine -> FillConvexPoly + Line2 + Circle) in integer arithmetic and
is PINNED against the cv2 build present in this image (4.13.0) by tests/test_raster_oracle.py: bit-exact on tens of
thousands of random segments including clipped and degenerate ones.  Version note: 4.13 first clips the segment to the image
rectangle grown by ``thickness`` pixels; older releases may differ for segments that leave the image by more than that
(face landmarks never do; shoulder points can).

Pure Python (small cases); the CUDA kernel (csrc/raster.cuh) implements the same integer algorithm one segment per thread.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT
THICKNESS = 2

# datasets/face_dataset.py:34-42
PART_LIST = [[list(range(0, 15))],
             [[15, 16, 17, 18, 18, 19, 20, 15]],
             [[21, 22, 23, 24, 24, 25, 26, 21]],
             [list(range(35, 44))],
             [[27, 65, 28, 68, 29], [29, 67, 30, 66, 27]],
             [[33, 69, 32, 72, 31], [31, 71, 34, 70, 33]],
             [list(range(46, 53)), [52, 53, 54, 55, 56, 57, 46]],
             [[46, 63, 62, 61, 52], [52, 60, 59, 58, 46]]]
N_LANDMARKS = 73


def face_segments() -> List[Tuple[int, int]]:
    """(landmark index, landmark index) of every cv2.line call of draw_face_feature_maps, in call order."""
    segs = []
    for edge_list in PART_LIST:
        for edge in edge_list:
            for i in range(len(edge) - 1):
                segs.append((edge[i], edge[i + 1]))
    return segs


def shoulder_segments(n_points: int) -> List[Tuple[int, int]]:
    """draw_shoulder_points (:300-309): two polylines over the halves of the point list."""
    num = n_points // 2
    return [(i * num + j, i * num + j + 1) for i in range(2) for j in range(num - 1)]


def _tdiv(a: int, b: int) -> int:
    """C integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def clip_line(w: int, h: int, p1, p2):
    """imgproc/drawing.cpp clipLine(Size2l, Point2l&, Point2l&): region codes, intersections computed in double and truncated."""
    x1, y1 = p1
    x2, y2 = p2
    right, bottom = w - 1, h - 1
    if w <= 0 or h <= 0:
        return False, p1, p2
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * float(x2 - x1) / float(y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * float(x2 - x1) / float(y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * float(y2 - y1) / float(x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * float(y2 - y1) / float(x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


def _put(img: np.ndarray, x: int, y: int) -> None:
    h, w = img.shape
    if 0 <= x < w and 0 <= y < h:
        img[y, x] = 255


def _line2(img: np.ndarray, pt1, pt2) -> None:
    """Line2: DDA between two 16.16 fixed-point points, one pixel per step along the major axis."""
    h, w = img.shape
    ok, pt1, pt2 = clip_line(w << XY_SHIFT, h << XY_SHIFT, pt1, pt2)
    if not ok:
        return
    x1, y1 = pt1
    x2, y2 = pt2
    dx, dy = x2 - x1, y2 - y1
    ax, ay = abs(dx), abs(dy)
    if ax > ay:
        if dx < 0:
            dy = -dy
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = XY_ONE, _tdiv(dy << XY_SHIFT, ax | 1)
        ecount = (x2 - x1) >> XY_SHIFT
    else:
        if dy < 0:
            dx = -dx
            x1, x2, y1, y2 = x2, x1, y2, y1
        x_step, y_step = _tdiv(dx << XY_SHIFT, ay | 1), XY_ONE
        ecount = (y2 - y1) >> XY_SHIFT
    x1 += XY_ONE >> 1
    y1 += XY_ONE >> 1
    _put(img, (x2 + (XY_ONE >> 1)) >> XY_SHIFT, (y2 + (XY_ONE >> 1)) >> XY_SHIFT)
    if ax > ay:
        x1 >>= XY_SHIFT
        while ecount >= 0:
            _put(img, x1, y1 >> XY_SHIFT)
            x1 += 1
            y1 += y_step
            ecount -= 1
    else:
        y1 >>= XY_SHIFT
        while ecount >= 0:
            _put(img, x1 >> XY_SHIFT, y1)
            x1 += x_step
            y1 += 1
            ecount -= 1


def _fill_convex_poly(img: np.ndarray, v: Sequence[Tuple[int, int]]) -> None:
    """FillConvexPoly(shift = XY_SHIFT, LINE_8): outline with Line2, then a two-edge scanline walk."""
    h, w = img.shape
    npts = len(v)
    delta = XY_ONE >> 1
    xmin = xmax = v[0][0]
    ymin = ymax = v[0][1]
    imin = 0
    p0 = v[npts - 1]
    for i in range(npts):
        p = v[i]
        if p[1] < ymin:
            ymin, imin = p[1], i
        ymax = max(ymax, p[1])
        xmax = max(xmax, p[0])
        xmin = min(xmin, p[0])
        _line2(img, p0, p)
        p0 = p
    xmin = (xmin + delta) >> XY_SHIFT
    xmax = (xmax + delta) >> XY_SHIFT
    ymin = (ymin + delta) >> XY_SHIFT
    ymax = (ymax + delta) >> XY_SHIFT
    if npts < 3 or xmax < 0 or ymax < 0 or xmin >= w or ymin >= h:
        return
    ymax = min(ymax, h - 1)
    ex = [-XY_ONE, -XY_ONE]
    edx = [0, 0]
    eye = [ymin, ymin]
    eidx = [imin, imin]
    edi = [1, npts - 1]
    y = ymin
    edges = npts
    while True:
        for i in range(2):
            if y >= eye[i]:
                idx0, di = eidx[i], edi[i]
                idx = idx0 + di
                if idx >= npts:
                    idx -= npts
                while True:                       # for (; edges-- > 0; )
                    more = edges > 0
                    edges -= 1
                    if not more:
                        break
                    ty = (v[idx][1] + delta) >> XY_SHIFT
                    if ty > y:
                        xs, xe = v[idx0][0], v[idx][0]
                        eye[i] = ty
                        edx[i] = _tdiv((xe - xs) * 2 + (ty - y), 2 * (ty - y))
                        ex[i] = xs
                        eidx[i] = idx
                        break
                    idx0 = idx
                    idx += di
                    if idx >= npts:
                        idx -= npts
        if edges < 0:
            break
        if y >= 0:
            left, right = (1, 0) if ex[0] > ex[1] else (0, 1)
            xx1 = (ex[left] + delta) >> XY_SHIFT
            xx2 = (ex[right] + delta) >> XY_SHIFT
            if xx2 >= 0 and xx1 < w:
                xx1 = max(xx1, 0)
                xx2 = min(xx2, w - 1)
                if xx2 >= xx1:
                    img[y, xx1:xx2 + 1] = 255
        ex[0] += edx[0]
        ex[1] += edx[1]
        y += 1
        if y > ymax:
            break


def _circle_r1(img: np.ndarray, cx: int, cy: int) -> None:
    """Circle(center, radius 1, filled): the 5-pixel plus shape (rows cy-1..cy+1)."""
    for x, y in ((cx - 1, cy), (cx, cy), (cx + 1, cy), (cx, cy - 1), (cx, cy + 1)):
        _put(img, x, y)


def thick_line(img: np.ndarray, pt1: Tuple[int, int], pt2: Tuple[int, int]) -> None:
    """cv2.line(img, pt1, pt2, 255, thickness=2) on a uint8 single-channel image, in place."""
    h, w = img.shape
    t = THICKNESS
    ok, q1, q2 = clip_line(w + 2 * t, h + 2 * t, (pt1[0] + t, pt1[1] + t), (pt2[0] + t, pt2[1] + t))
    if not ok:
        return
    p0 = ((q1[0] - t) << XY_SHIFT, (q1[1] - t) << XY_SHIFT)
    p1 = ((q2[0] - t) << XY_SHIFT, (q2[1] - t) << XY_SHIFT)
    dx = (p0[0] - p1[0]) * (1.0 / XY_ONE)
    dy = (p1[1] - p0[1]) * (1.0 / XY_ONE)
    r = dx * dx + dy * dy
    half = t << (XY_SHIFT - 1)
    if abs(r) > np.finfo(np.float64).eps:
        r = half / math.sqrt(r)
        dpx = int(np.rint(dy * r))             # cvRound: round half to even
        dpy = int(np.rint(dx * r))
        _fill_convex_poly(img, [(p0[0] + dpx, p0[1] + dpy), (p0[0] - dpx, p0[1] - dpy),
                                (p1[0] - dpx, p1[1] - dpy), (p1[0] + dpx, p1[1] + dpy)])
    for p in (p0, p1):
        _circle_r1(img, (p[0] + (XY_ONE >> 1)) >> XY_SHIFT, (p[1] + (XY_ONE >> 1)) >> XY_SHIFT)


def _trunc_points(pts: np.ndarray) -> List[Tuple[int, int]]:
    return [(int(p[0]), int(p[1])) for p in pts]       # Python int(): toward zero, as the reference does


def draw_feature_map(landmarks: np.ndarray, size: Tuple[int, int] = (512, 512), shoulders: Optional[np.ndarray] = None) -> np.ndarray:
    """get_feature_image (:285-297) without the image_pad shift: uint8 [H,W] in {0,255}."""
    w, h = size
    img = np.zeros((h, w), np.uint8)
    lm = _trunc_points(landmarks)
    for a, b in face_segments():
        thick_line(img, lm[a], lm[b])
    if shoulders is not None:
        sp = _trunc_points(shoulders)
        for a, b in shoulder_segments(len(sp)):
            thick_line(img, sp[a], sp[b])
    return img


def feature_map_tensor(landmarks: np.ndarray, size=(512, 512), shoulders: Optional[np.ndarray] = None) -> np.ndarray:
    """get_data_test_mode (:276-282): float32 [1,H,W] in {0,1}."""
    return draw_feature_map(landmarks, size, shoulders)[None].astype(np.float32) / 255.0


# ---------------------------------------------------------------------------------------------------------------
# The same drawing through the reference's own dependency (cv2), used to pin the restatement and as the CPU baseline.
# ---------------------------------------------------------------------------------------------------------------
def draw_feature_map_cv2(landmarks: np.ndarray, size=(512, 512), shoulders: Optional[np.ndarray] = None) -> np.ndarray:
    import cv2

    w, h = size
    img = np.zeros((h, w), np.uint8)
    lm = _trunc_points(landmarks)
    for a, b in face_segments():
        img = cv2.line(img, lm[a], lm[b], 255, 2)
    if shoulders is not None:
        sp = _trunc_points(shoulders)
        for a, b in shoulder_segments(len(sp)):
            img = cv2.line(img, sp[a], sp[b], 255, 2)
    return img


def make_landmarks(batch: int, size=(512, 512), seed: int = 3, spill: float = 0.0):
    """Seeded synthetic landmark tracks: 73 points on noisy closed curves around a face centre (consecutive indices are
    neighbours, as in the tracked data, so segments are 5-40 px long) and 18 shoulder points along the bottom edge, some of
    them outside the image as in the reference data; ``spill`` > 0 scales everything up to push points across the borders
    and exercise the clipping paths."""
    rng = np.random.default_rng(seed)
    w, h = size
    i = np.arange(N_LANDMARKS)
    ang = 2 * np.pi * i / 18.0
    rad = w * (0.08 + 0.12 * ((i * 7) % 5) / 4.0) * (1.0 + 3.0 * spill)
    base = np.stack([np.cos(ang) * rad, np.sin(ang) * rad * 1.2], 1)
    centre = np.array([w / 2, h * 0.45]) + rng.normal(0, 6, (batch, 1, 2))
    lm = centre + base[None] + rng.normal(0, 2.5, (batch, N_LANDMARKS, 2))
    xs = np.linspace(-0.05 - spill, 1.05 + spill, 9) * w
    sh = np.stack([np.concatenate([xs, xs]), np.concatenate([np.full(9, h * 0.93), np.full(9, h * 1.01)])], 1)[None]
    sh = sh + rng.normal(0, 4, (batch, 18, 2))
    return lm.astype(np.float32), sh.astype(np.float32)
