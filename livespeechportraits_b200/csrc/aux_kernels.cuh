// Memory-bound helper kernels around the conv stack.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace lspg {

// Input packer: fuses Feature2FaceModel.inference's torch.cat([feature_map, cand_image], 1)
// (reference models/feature2face_model.py:231), the NCHW->NHWC transpose, the fp32->16-bit conversion (bf16 in FAST mode,
// fp16 hi + lo limbs in PARITY mode) and a 2x2 space-to-depth, so that the stride-2 head conv (13->64,
// models/networks.py:594-595 with input_nc=13) becomes a 4-tap stride-1 conv over a 64-channel tensor:
//   S[n, oy, ox, (py*2+px)*16 + c] = x[n, c, 2*oy+py, 2*ox+px]   (c < in_nc; channels in_nc..15 are zero)
// One thread per output pixel, one warp per 32 consecutive pixels of an output row: the float2 reads are coalesced along W
// per channel plane (256 B per warp), and the 32 NHWC rows of a warp (32 x 128 B per limb = 4 KB contiguous) go through a
// swizzled shared-memory transpose so that every store instruction of the warp writes 512 contiguous bytes.
// HBM-bound: 4*in_nc*H*W bytes in, 2*NL*16*H*W bytes out per frame.  Requires W/2 to be a multiple of 32 (W % 256 == 0).
constexpr int kPackWarps = 4;
template <int NL>
__global__ void __launch_bounds__(kPackWarps * 32) pack_input_s2d_kernel(const float* __restrict__ fm, long long fm_bstride,
                                                                       const float* __restrict__ cand, long long cand_bstride,
                                                                       int in_nc, __nv_bfloat16* __restrict__ dst,
                                                                       long long limb_stride, int batch, int height, int width) {
  __shared__ uint4 stage[kPackWarps][NL][32][8];
  // PDL: the previous forward's kernels may still be reading the packed tensor this kernel overwrites
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int wo = width >> 1, ho = height >> 1;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long groups = static_cast<long long>(batch) * ho * (wo >> 5);        // groups of 32 pixels
  const long long plane = static_cast<long long>(height) * width;
  for (long long g = blockIdx.x * static_cast<long long>(kPackWarps) + wib; g < groups;
       g += static_cast<long long>(gridDim.x) * kPackWarps) {
    const long long i0 = g << 5;                 // first output pixel of the group
    const long long i = i0 + lane;
    const int ox = static_cast<int>(i % wo);
    const long long r = i / wo;
    const int oy = static_cast<int>(r % ho);
    const int n = static_cast<int>(r / ho);
    uint32_t hi[32], lo[32];   // 64 channels packed as bf16 pairs
#pragma unroll
    for (int k = 0; k < 32; ++k) { hi[k] = 0u; lo[k] = 0u; }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      if (c < in_nc) {
        const float* src = (c == 0) ? fm + n * fm_bstride : cand + n * cand_bstride + (c - 1) * plane;
#pragma unroll
        for (int py = 0; py < 2; ++py) {
          const float2 v = *reinterpret_cast<const float2*>(src + static_cast<long long>(2 * oy + py) * width + 2 * ox);
          const float vv[2] = {v.x, v.y};
#pragma unroll
          for (int px = 0; px < 2; ++px) {
            const int ch = (py * 2 + px) * 16 + c;
            if (NL == 2) {                      // PARITY: fp16 hi + lo limbs (inputs are in [-1, 1]: no range concern)
              const __half h = __float2half_rn(vv[px]);
              hi[ch >> 1] |= static_cast<uint32_t>(__half_as_ushort(h)) << ((ch & 1) * 16);
              const __half l = __float2half_rn(vv[px] - __half2float(h));
              lo[ch >> 1] |= static_cast<uint32_t>(__half_as_ushort(l)) << ((ch & 1) * 16);
            } else {                            // FAST: bf16
              const __nv_bfloat16 h = __float2bfloat16_rn(vv[px]);
              hi[ch >> 1] |= static_cast<uint32_t>(__bfloat16_as_ushort(h)) << ((ch & 1) * 16);
            }
          }
        }
      }
    }
    __syncwarp();                                // the previous group's reads of the staging rows are done
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      stage[wib][0][lane][k ^ (lane & 7)] = make_uint4(hi[4 * k], hi[4 * k + 1], hi[4 * k + 2], hi[4 * k + 3]);
      if (NL == 2) stage[wib][NL - 1][lane][k ^ (lane & 7)] = make_uint4(lo[4 * k], lo[4 * k + 1], lo[4 * k + 2], lo[4 * k + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      uint4* d = reinterpret_cast<uint4*>(dst + l * limb_stride + i0 * 64);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int q = j * 32 + lane, pp = q >> 3, kk = q & 7;
        d[q] = stage[wib][l][pp][kk ^ (pp & 7)];
      }
    }
  }
}

}  // namespace lspg
