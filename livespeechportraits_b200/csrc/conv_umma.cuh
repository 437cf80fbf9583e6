// Fused 3x3 convolution as an implicit GEMM on tcgen05 tensor cores (sm_100a).
//
// One kernel family serves every conv of the Res-UNet generator
// (reference: models/networks.py:585-675):
//
//   D[128 pixels x BN channels] (+)= A[128 pixels x 64 ch] * B[BN x 64 ch]^T      per K block
//
//   * A K block = (tap, concat source, 64-channel chunk).  The A tile of a tap is the output tile's pixel
//     patch shifted by (dx, dy): ONE 5-D TMA box {64 ch, TW, TH, NB, 1 limb} out of the NHWC activation
//     tensor; out-of-range coordinates are zero-filled by TMA, which is the conv's zero padding.
//   * stride-2 convs read through four parity views of the input (base offset + doubled strides), the
//     nearest-x2 upsample is folded into four output phases with 2x2 pre-summed taps whose results are
//     stored through four phase views of the output, and torch.cat([skip, deeper]) is two K ranges fed from
//     two tensor maps - so upsample, concat, padding and stride never touch HBM as separate passes.
//   * B tiles come from a packed weight tensor [phase*limb][Cout][K] (K-major), same 128B-swizzled layout.
//   * Accumulators live in TMEM (double buffered, 2 x BN columns); the epilogue reads them with
//     tcgen05.ld, applies the folded eval-BatchNorm scale/shift in fp32, adds the residual tile (TMA-loaded
//     into the same staging buffer the result is written back to), ReLU, converts to bf16 and hands the
//     tile to a TMA store.  The tail variant applies tanh and scatters fp32 NCHW directly.
//   * NL = 2 ("parity" precision): activations and weights are split bf16 hi + lo limbs; each K step issues
//     hi*hi + hi*lo + lo*hi (~16 mantissa bits, fp32 accumulate) and the epilogue writes both limbs.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = epilogue
// (TMEM lane quarter = warp_id % 4).  Persistent: each CTA walks tiles blockIdx.x, +gridDim.x, ...
#pragma once
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace lspg {

constexpr int kTileM = 128;                 // output pixels per tile (UMMA M)
constexpr int kChunk = 64;                  // channels per K block = 128 bytes of bf16 = one swizzle row
constexpr int kATile = kTileM * 128;        // bytes of one A (or staging) tile
constexpr int kMaxTaps = 9;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kSmemBudget = 227 * 1024;

struct alignas(64) ConvParams {
  CUtensorMap a[4];          // activation views: [parity or concat source], dims {C, X, Y, N, limb}
  CUtensorMap w;             // packed weights, dims {K, Cout_pad, limb*n_phases + phase}
  CUtensorMap out[4];        // output views per phase, dims {C, X, Y, N, limb}
  CUtensorMap res;           // residual view
  const float* scale;        // folded BatchNorm (or 1/0), [Cout_pad]
  const float* shift;
  float* out_f32;            // tail only: fp32 NCHW [B, 3, 2*hs, 2*ws]
  uint32_t idesc;            // UMMA instruction descriptor (M=128, N=BN, bf16 x bf16 -> f32, K-major)
  int32_t n_taps, n_src;
  int32_t chunks[2];         // 64-channel chunks per concat source
  int32_t tiles_x, tiles_y, tiles_n;
  int32_t tw_log2, th_log2;  // tile = TW x TH pixels x NB images, TW*TH*NB = 128
  int32_t n_tiles;           // Cout_pad / BN
  int32_t n_phases;
  int32_t total_tiles;       // tiles_x*tiles_y*tiles_n * n_tiles * n_phases
  int32_t relu, has_res;
  int32_t batch, hs, ws;     // sampling-grid extent (tail epilogue addressing)
  int8_t tap_map[4][kMaxTaps];   // [phase][tap] -> index into a[] (added to the concat source index)
  int8_t tap_dx[4][kMaxTaps];
  int8_t tap_dy[4][kMaxTaps];
};

template <int BN, int NL, bool TAIL>
struct ConvCfg {
  static constexpr int kBTile = BN * 128;
  static constexpr int kStage = NL * (kATile + kBTile);
  static constexpr int kNumStg = TAIL ? 0 : (NL == 1 ? 2 : 1);       // residual-in / result-out staging buffers
  static constexpr int kStgBytes = kNumStg * NL * kATile;
  static constexpr int kAux = 2048;                                  // scale/shift + barriers + tmem ptr
  static constexpr int kAvail = kSmemBudget - 1024 /*alignment slack*/ - kStgBytes - kAux;
  static constexpr int kStagesRaw = kAvail / kStage;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = 1024 + kStages * kStage + kStgBytes + kAux;
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;
  static_assert(kStages >= 2, "pipeline needs at least two stages");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");
  static_assert(2 * BN * 4 + 256 <= kAux, "aux region too small");
};

struct TileCoord {
  int z, nt, x0, y0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t) {
  TileCoord c;
  const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
  int mt = t % m_tiles;
  int r = t / m_tiles;
  c.nt = r % p.n_tiles;
  c.z = r / p.n_tiles;
  int tx = mt % p.tiles_x;
  int r2 = mt / p.tiles_x;
  int ty = r2 % p.tiles_y;
  int tn = r2 / p.tiles_y;
  c.x0 = tx << p.tw_log2;
  c.y0 = ty << p.th_log2;
  c.n0 = tn << (7 - p.tw_log2 - p.th_log2);
  return c;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

template <int BN, int NL, bool TAIL>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, NL, TAIL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + Cfg::kStages * Cfg::kStage;
  uint8_t* aux = stg_base + Cfg::kStgBytes;
  float* s_scale = reinterpret_cast<float*>(aux);
  float* s_shift = s_scale + BN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux + 2 * BN * 4);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint64_t* stg_bar = tempty_bar + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.a[i]);
    ptx::prefetch_tmap(&p.w);
    if (!TAIL) {
      for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&p.out[i]);
      ptx::prefetch_tmap(&p.res);
    }
    for (int i = 0; i < Cfg::kStages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 4);   // one arrive per epilogue warp
      ptx::mbar_init(&stg_bar[i], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int kb_per_tap = 0;
  for (int s = 0; s < p.n_src; ++s) kb_per_tap += p.chunks[s];
  const int num_kb = p.n_taps * kb_per_tap;

  if (warp == 0) {
    // ===================================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t);
      int kb = 0;
      for (int tap = 0; tap < p.n_taps; ++tap) {
        const int amap = p.tap_map[tc.z][tap];
        const int xx = tc.x0 + p.tap_dx[tc.z][tap];
        const int yy = tc.y0 + p.tap_dy[tc.z][tap];
        for (int s = 0; s < p.n_src; ++s) {
          for (int c = 0; c < p.chunks[s]; ++c, ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            if (lane == 0) {
              uint8_t* st = smem + stage * Cfg::kStage;
              ptx::mbar_expect_tx(&full_bar[stage], Cfg::kStage);
#pragma unroll
              for (int l = 0; l < NL; ++l) {
                ptx::tma_load_5d(&p.a[amap + s], &full_bar[stage], st + l * kATile, c * kChunk, xx, yy, tc.n0, l);
                ptx::tma_load_3d(&p.w, &full_bar[stage], st + NL * kATile + l * Cfg::kBTile, kb * kChunk, tc.nt * BN,
                                 l * p.n_phases + tc.z);
              }
            }
            __syncwarp();
            if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (lane == 0) {
          const uint32_t a0 = ptx::smem_u32(smem + stage * Cfg::kStage);
          const uint32_t b0 = a0 + NL * kATile;
          const uint64_t a_hi = ptx::umma_desc_sw128(a0);
          const uint64_t b_hi = ptx::umma_desc_sw128(b0);
#pragma unroll
          for (int k = 0; k < kChunk / 16; ++k) {
            const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
            // +32 bytes along K inside the 128-byte swizzle row = +2 in descriptor address units
            ptx::umma_f16(d_tmem, a_hi + 2 * k, b_hi + 2 * k, p.idesc, accum);
            if (NL == 2) {
              const uint64_t a_lo = ptx::umma_desc_sw128(a0 + kATile);
              const uint64_t b_lo = ptx::umma_desc_sw128(b0 + Cfg::kBTile);
              ptx::umma_f16(d_tmem, a_hi + 2 * k, b_lo + 2 * k, p.idesc, 1u);
              ptx::umma_f16(d_tmem, a_lo + 2 * k, b_hi + 2 * k, p.idesc, 1u);
            }
          }
          ptx::umma_commit(&empty_bar[stage]);                 // smem slot reusable once these MMAs retire
          if (kb == num_kb - 1) ptx::umma_commit(&tfull_bar[acc]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    const int q = warp & 3;                     // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;              // tile row = pixel index inside the tile
    const int etid = threadIdx.x - 64;          // 0..127
    const bool leader = (etid == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t g = 0;                             // running staging-chunk counter (selects buffer + parity)

    // The leader "prepares" staging buffer g % kNumStg for chunk g: waits until the TMA store that last read
    // it has drained, then either TMA-loads the residual chunk into it or simply releases it.
    auto prepare = [&](uint32_t gg, const TileCoord& tc, int chunk) {
      if (TAIL) return;
      const int j = gg % (Cfg::kNumStg > 0 ? Cfg::kNumStg : 1);
      uint8_t* buf = stg_base + j * NL * kATile;
      if (p.has_res) {
        ptx::mbar_expect_tx(&stg_bar[j], NL * kATile);
#pragma unroll
        for (int l = 0; l < NL; ++l)
          ptx::tma_load_5d(&p.res, &stg_bar[j], buf + l * kATile, tc.nt * BN + chunk * kChunk, tc.x0, tc.y0, tc.n0, l);
      } else {
        ptx::mbar_arrive(&stg_bar[j]);
      }
    };

    constexpr int kChunksPerTile = TAIL ? 1 : BN / kChunk;
    bool first = true;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(p, t);
      // folded BatchNorm parameters of this tile's channel range
      for (int i = etid; i < BN; i += kEpiThreads) {
        s_scale[i] = p.scale[tc.nt * BN + i];
        s_shift[i] = p.shift[tc.nt * BN + i];
      }
      if (!TAIL && Cfg::kNumStg == 2 && first && leader) prepare(g, tc, 0);
      first = false;
      ptx::named_bar_sync(1, kEpiThreads);

      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;

      if constexpr (TAIL) {
        // ---- tail: 16 columns = 4 phases x 3 channels (+4 pad); tanh; fp32 NCHW scatter
        uint32_t v[16];
        ptx::tmem_ld_32x16(t_acc, v);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        const int tw = row & ((1 << p.tw_log2) - 1);
        const int th = (row >> p.tw_log2) & ((1 << p.th_log2) - 1);
        const int nb = row >> (p.tw_log2 + p.th_log2);
        const int n = tc.n0 + nb, y = tc.y0 + th, x = tc.x0 + tw;
        if (n < p.batch && y < p.hs && x < p.ws) {
          const int oh = 2 * p.hs, ow = 2 * p.ws;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int py = 0; py < 2; ++py) {
              float2 o;
              o.x = tanhf(__uint_as_float(v[(py * 2 + 0) * 3 + c]) * s_scale[(py * 2 + 0) * 3 + c] +
                          s_shift[(py * 2 + 0) * 3 + c]);
              o.y = tanhf(__uint_as_float(v[(py * 2 + 1) * 3 + c]) * s_scale[(py * 2 + 1) * 3 + c] +
                          s_shift[(py * 2 + 1) * 3 + c]);
              float* dst = p.out_f32 + ((static_cast<size_t>(n) * 3 + c) * oh + (2 * y + py)) * ow + 2 * x;
              *reinterpret_cast<float2*>(dst) = o;
            }
          }
        }
      } else {
        for (int chunk = 0; chunk < kChunksPerTile; ++chunk, ++g) {
          const int j = g % Cfg::kNumStg;
          const uint32_t par = (g / Cfg::kNumStg) & 1;
          uint8_t* buf = stg_base + j * NL * kATile;
          if (leader) {
            // all earlier stores must have finished READING their staging buffer before it is refilled
            ptx::tma_store_wait_read<0>();
            if (Cfg::kNumStg == 2) {
              // prefetch the next chunk's residual into the other buffer (its last store has drained)
              int nchunk = chunk + 1;
              int nt_tile = t;
              if (nchunk == kChunksPerTile) { nchunk = 0; nt_tile = t + gridDim.x; }
              if (nt_tile < p.total_tiles) {
                const TileCoord ntc = (nt_tile == t) ? tc : decode_tile(p, nt_tile);
                prepare(g + 1, ntc, nchunk);
              }
            } else {
              prepare(g, tc, chunk);
            }
          }
          ptx::mbar_wait(&stg_bar[j], par);

          uint8_t* my_row = buf + row * 128;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(t_acc + chunk * kChunk + half * 32, v);
            ptx::tmem_ld_wait();
            if (chunk == kChunksPerTile - 1 && half == 1) {
              // accumulator fully drained into registers: hand the TMEM buffer back to the MMA warp
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
            }
#pragma unroll
            for (int c16 = 0; c16 < 4; ++c16) {            // 16-byte pieces: 8 channels each
              const int piece = half * 4 + c16;            // logical 16B chunk index inside the 128B row
              const int phys = (piece ^ (row & 7)) << 4;   // 128B swizzle
              float y[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int ch = chunk * kChunk + piece * 8 + e;
                y[e] = fmaf(__uint_as_float(v[c16 * 8 + e]), s_scale[ch], s_shift[ch]);
              }
              if (p.has_res) {
                const uint4 r = *reinterpret_cast<const uint4*>(my_row + phys);
                const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  y[2 * e] += bf16_lo(rr[e]);
                  y[2 * e + 1] += bf16_hi(rr[e]);
                }
                if (NL == 2) {
                  const uint4 r2 = *reinterpret_cast<const uint4*>(my_row + kATile + phys);
                  const uint32_t rr2[4] = {r2.x, r2.y, r2.z, r2.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    y[2 * e] += bf16_lo(rr2[e]);
                    y[2 * e + 1] += bf16_hi(rr2[e]);
                  }
                }
              }
              if (p.relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], 0.0f);
              }
              uint4 o;
              o.x = pack_bf16x2(y[0], y[1]);
              o.y = pack_bf16x2(y[2], y[3]);
              o.z = pack_bf16x2(y[4], y[5]);
              o.w = pack_bf16x2(y[6], y[7]);
              *reinterpret_cast<uint4*>(my_row + phys) = o;
              if (NL == 2) {
                const uint32_t oo[4] = {o.x, o.y, o.z, o.w};
                uint32_t lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  lo[e] = pack_bf16x2(y[2 * e] - bf16_lo(oo[e]), y[2 * e + 1] - bf16_hi(oo[e]));
                *reinterpret_cast<uint4*>(my_row + kATile + phys) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              }
            }
          }
          ptx::fence_proxy_async_smem();
          ptx::named_bar_sync(2, kEpiThreads);
          if (leader) {
#pragma unroll
            for (int l = 0; l < NL; ++l)
              ptx::tma_store_5d(&p.out[tc.z], buf + l * kATile, tc.nt * BN + chunk * kChunk, tc.x0, tc.y0, tc.n0, l);
            ptx::tma_store_commit();
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (!TAIL && leader) ptx::tma_store_wait_all<0>();   // global writes complete before the CTA retires
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace lspg
