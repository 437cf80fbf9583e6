// Feature-map rasteriser: landmark tracks -> the [B,1,H,W] {0,1} maps the generator consumes.
//
// Replaces, for a whole clip at once, the per-frame host loop of the reference
//   datasets/face_dataset.py:312-323 draw_face_feature_maps   (72 cv2.line calls, colour 255, thickness 2)
//   datasets/face_dataset.py:300-309 draw_shoulder_points     (2 polylines)
//   datasets/face_dataset.py:276-282 get_data_test_mode       (uint8 -> float32 / 255)
// and the 1 MB host->device copy per frame that follows it (demo.py:262-265).
//
// All segments have the same colour, so the map is the union of the pixel sets of the individual cv2.line calls and the
// calls are independent: one thread rasterises one segment with the integer algorithm of OpenCV's thick line
// (clip to the image grown by the thickness; ThickLine = FillConvexPoly of the 4-corner polygon in 16.16 fixed point,
// whose outline is drawn with Line2, plus a radius-1 filled circle at both ends) and scatters 1.0f into a zeroed map.
// Bit-exact with cv2 4.13 (oracle/raster_oracle.py restates the same algorithm and is pinned against cv2).
// Integer / byte work, a few hundred stores per thread: latency-trivial next to the generator (about 10 us per clip batch).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lspg {

constexpr int kRasterShift = 16;
constexpr long long kRasterOne = 1ll << kRasterShift;
constexpr int kRasterThickness = 2;
constexpr int kFaceSegments = 72;
constexpr int kFaceLandmarks = 73;

// consecutive pairs of datasets/face_dataset.py:34-42 part_list, in call order
__constant__ uint8_t c_face_seg[kFaceSegments][2] = {
    {0, 1},   {1, 2},   {2, 3},   {3, 4},   {4, 5},   {5, 6},   {6, 7},   {7, 8},   {8, 9},   {9, 10},  {10, 11}, {11, 12},
    {12, 13}, {13, 14},                                                                                     // contour
    {15, 16}, {16, 17}, {17, 18}, {18, 18}, {18, 19}, {19, 20}, {20, 15},                                   // right eyebrow
    {21, 22}, {22, 23}, {23, 24}, {24, 24}, {24, 25}, {25, 26}, {26, 21},                                   // left eyebrow
    {35, 36}, {36, 37}, {37, 38}, {38, 39}, {39, 40}, {40, 41}, {41, 42}, {42, 43},                         // nose
    {27, 65}, {65, 28}, {28, 68}, {68, 29}, {29, 67}, {67, 30}, {30, 66}, {66, 27},                         // right eye
    {33, 69}, {69, 32}, {32, 72}, {72, 31}, {31, 71}, {71, 34}, {34, 70}, {70, 33},                         // left eye
    {46, 47}, {47, 48}, {48, 49}, {49, 50}, {50, 51}, {51, 52}, {52, 53}, {53, 54}, {54, 55}, {55, 56}, {56, 57}, {57, 46},  // mouth
    {46, 63}, {63, 62}, {62, 61}, {61, 52}, {52, 60}, {60, 59}, {59, 58}, {58, 46}};                        // tongue

struct RPoint {
  long long x, y;
};

struct RasterTarget {
  float* img;      // one frame, [H][W]
  int w, h;
  __device__ __forceinline__ void put(long long x, long long y) const {
    if (x >= 0 && x < w && y >= 0 && y < h) img[y * w + x] = 1.0f;
  }
  __device__ __forceinline__ void hline(int y, int x0, int x1) const {      // clipped by the caller
    float* row = img + static_cast<long long>(y) * w;
    for (int x = x0; x <= x1; ++x) row[x] = 1.0f;
  }
};

// Cohen-Sutherland clip of a segment to [0,w) x [0,h); the intersections are computed in double and truncated.
__device__ inline bool raster_clip(long long w, long long h, RPoint& p1, RPoint& p2) {
  const long long right = w - 1, bottom = h - 1;
  if (w <= 0 || h <= 0) return false;
  long long &x1 = p1.x, &y1 = p1.y, &x2 = p2.x, &y2 = p2.y;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += static_cast<long long>(__ddiv_rn(__dmul_rn(static_cast<double>(a - y1), static_cast<double>(x2 - x1)), static_cast<double>(y2 - y1)));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += static_cast<long long>(__ddiv_rn(__dmul_rn(static_cast<double>(a - y2), static_cast<double>(x2 - x1)), static_cast<double>(y2 - y1)));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += static_cast<long long>(__ddiv_rn(__dmul_rn(static_cast<double>(a - x1), static_cast<double>(y2 - y1)), static_cast<double>(x2 - x1)));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += static_cast<long long>(__ddiv_rn(__dmul_rn(static_cast<double>(a - x2), static_cast<double>(y2 - y1)), static_cast<double>(x2 - x1)));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

// One-pixel-wide DDA between two 16.16 points (the polygon outline).
__device__ inline void raster_line2(const RasterTarget& t, RPoint p1, RPoint p2) {
  if (!raster_clip(static_cast<long long>(t.w) << kRasterShift, static_cast<long long>(t.h) << kRasterShift, p1, p2)) return;
  long long dx = p2.x - p1.x, dy = p2.y - p1.y;
  const long long ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
  long long x_step, y_step;
  int ecount;
  if (ax > ay) {
    if (dx < 0) {
      dy = -dy;
      const RPoint s = p1; p1 = p2; p2 = s;
    }
    x_step = kRasterOne;
    y_step = (dy << kRasterShift) / (ax | 1);
    ecount = static_cast<int>((p2.x - p1.x) >> kRasterShift);
  } else {
    if (dy < 0) {
      dx = -dx;
      const RPoint s = p1; p1 = p2; p2 = s;
    }
    x_step = (dx << kRasterShift) / (ay | 1);
    y_step = kRasterOne;
    ecount = static_cast<int>((p2.y - p1.y) >> kRasterShift);
  }
  p1.x += kRasterOne >> 1;
  p1.y += kRasterOne >> 1;
  t.put((p2.x + (kRasterOne >> 1)) >> kRasterShift, (p2.y + (kRasterOne >> 1)) >> kRasterShift);
  if (ax > ay) {
    p1.x >>= kRasterShift;
    for (; ecount >= 0; --ecount) {
      t.put(p1.x, p1.y >> kRasterShift);
      p1.x += 1;
      p1.y += y_step;
    }
  } else {
    p1.y >>= kRasterShift;
    for (; ecount >= 0; --ecount) {
      t.put(p1.x >> kRasterShift, p1.y);
      p1.x += x_step;
      p1.y += 1;
    }
  }
}

// Convex polygon of 4 points in 16.16 fixed point: outline, then a scanline walk along the left and right edge chains.
__device__ inline void raster_fill_quad(const RasterTarget& t, const RPoint (&v)[4]) {
  constexpr int npts = 4;
  const long long delta = kRasterOne >> 1;
  long long xmin = v[0].x, xmax = v[0].x, ymin = v[0].y, ymax = v[0].y;
  int imin = 0;
  RPoint p0 = v[npts - 1];
#pragma unroll
  for (int i = 0; i < npts; ++i) {
    const RPoint p = v[i];
    if (p.y < ymin) { ymin = p.y; imin = i; }
    ymax = p.y > ymax ? p.y : ymax;
    xmax = p.x > xmax ? p.x : xmax;
    xmin = p.x < xmin ? p.x : xmin;
    raster_line2(t, p0, p);
    p0 = p;
  }
  xmin = (xmin + delta) >> kRasterShift;
  xmax = (xmax + delta) >> kRasterShift;
  ymin = (ymin + delta) >> kRasterShift;
  ymax = (ymax + delta) >> kRasterShift;
  if (xmax < 0 || ymax < 0 || xmin >= t.w || ymin >= t.h) return;
  if (ymax > t.h - 1) ymax = t.h - 1;
  long long ex[2] = {-kRasterOne, -kRasterOne}, edx[2] = {0, 0};
  int eye[2] = {static_cast<int>(ymin), static_cast<int>(ymin)};
  int eidx[2] = {imin, imin};
  const int edi[2] = {1, npts - 1};
  int y = static_cast<int>(ymin);
  int edges = npts;
  do {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (y >= eye[i]) {
        int idx0 = eidx[i];
        const int di = edi[i];
        int idx = idx0 + di;
        if (idx >= npts) idx -= npts;
        for (; edges-- > 0;) {
          const int ty = static_cast<int>((v[idx].y + delta) >> kRasterShift);
          if (ty > y) {
            const long long xs = v[idx0].x, xe = v[idx].x;
            eye[i] = ty;
            edx[i] = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
            ex[i] = xs;
            eidx[i] = idx;
            break;
          }
          idx0 = idx;
          idx += di;
          if (idx >= npts) idx -= npts;
        }
      }
    }
    if (edges < 0) break;
    if (y >= 0) {
      const int left = ex[0] > ex[1] ? 1 : 0, right = left ^ 1;
      long long xx1 = (ex[left] + delta) >> kRasterShift;
      long long xx2 = (ex[right] + delta) >> kRasterShift;
      if (xx2 >= 0 && xx1 < t.w) {
        if (xx1 < 0) xx1 = 0;
        if (xx2 >= t.w) xx2 = t.w - 1;
        t.hline(y, static_cast<int>(xx1), static_cast<int>(xx2));
      }
    }
    ex[0] += edx[0];
    ex[1] += edx[1];
  } while (++y <= static_cast<int>(ymax));
}

// cv2.line(img, pt1, pt2, 255, thickness = 2) restricted to "set covered pixels to 1.0f".
__device__ inline void raster_thick_line(const RasterTarget& t, long long x1, long long y1, long long x2, long long y2) {
  constexpr int th = kRasterThickness;
  RPoint q1 = {x1 + th, y1 + th}, q2 = {x2 + th, y2 + th};
  if (!raster_clip(t.w + 2 * th, t.h + 2 * th, q1, q2)) return;
  RPoint p0 = {(q1.x - th) << kRasterShift, (q1.y - th) << kRasterShift};
  RPoint p1 = {(q2.x - th) << kRasterShift, (q2.y - th) << kRasterShift};
  const double inv = 1.0 / static_cast<double>(kRasterOne);
  const double dx = static_cast<double>(p0.x - p1.x) * inv, dy = static_cast<double>(p1.y - p0.y) * inv;   // exact (integers)
  double r = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
  const long long half = static_cast<long long>(th) << (kRasterShift - 1);
  if (fabs(r) > 2.220446049250313e-16) {
    r = __ddiv_rn(static_cast<double>(half), __dsqrt_rn(r));
    const long long dpx = __double2ll_rn(__dmul_rn(dy, r));     // cvRound: round half to even
    const long long dpy = __double2ll_rn(__dmul_rn(dx, r));
    const RPoint v[4] = {{p0.x + dpx, p0.y + dpy}, {p0.x - dpx, p0.y - dpy}, {p1.x - dpx, p1.y - dpy}, {p1.x + dpx, p1.y + dpy}};
    raster_fill_quad(t, v);
  }
#pragma unroll
  for (int e = 0; e < 2; ++e) {      // filled circle of radius 1 at both ends: the 5-pixel plus shape
    const RPoint p = e ? p1 : p0;
    const long long cx = (p.x + (kRasterOne >> 1)) >> kRasterShift, cy = (p.y + (kRasterOne >> 1)) >> kRasterShift;
    t.put(cx - 1, cy); t.put(cx, cy); t.put(cx + 1, cy); t.put(cx, cy - 1); t.put(cx, cy + 1);
  }
}

// grid = (ceil(segments / blockDim.x), B); the maps must be zero before the launch.
__global__ void raster_feature_maps_kernel(const float* __restrict__ landmarks, const float* __restrict__ shoulders, int n_shoulder,
                                           float* __restrict__ out, int H, int W) {
  const int seg = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  const int half_n = n_shoulder / 2;
  const int n_sh_seg = (shoulders != nullptr && half_n > 1) ? 2 * (half_n - 1) : 0;
  if (seg >= kFaceSegments + n_sh_seg) return;
  const float* pts;
  int ia, ib;
  if (seg < kFaceSegments) {
    pts = landmarks + static_cast<size_t>(b) * kFaceLandmarks * 2;
    ia = c_face_seg[seg][0];
    ib = c_face_seg[seg][1];
  } else {
    const int s = seg - kFaceSegments;
    const int side = s / (half_n - 1), j = s - side * (half_n - 1);
    pts = shoulders + static_cast<size_t>(b) * n_shoulder * 2;
    ia = side * half_n + j;
    ib = ia + 1;
  }
  // Python int(float): truncation toward zero
  const long long x1 = static_cast<long long>(pts[2 * ia]), y1 = static_cast<long long>(pts[2 * ia + 1]);
  const long long x2 = static_cast<long long>(pts[2 * ib]), y2 = static_cast<long long>(pts[2 * ib + 1]);
  const RasterTarget t = {out + static_cast<size_t>(b) * H * W, W, H};
  raster_thick_line(t, x1, y1, x2, y2);
}

}  // namespace lspg
