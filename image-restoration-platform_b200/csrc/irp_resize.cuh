// irp_resize.cuh — preprocess pixel stages (sm_100a):
//   P2 EXIF auto-orient  (reference: server-node/src/middleware/imagePreprocess.js:42, sharp .rotate())
//   P3 lanczos3 fit-inside resize = libvips reducev then reduceh, 12-bit fixed point,
//      u8 between passes                                   (imagePreprocess.js:46-53)
//   P4 normalise to raw u8 (RGBA flattened on black)       (imagePreprocess.js:57-63, pixel part)
//   P5 placement on a 2048x2048 fusion canvas              
//
// One CTA produces one output tile.  The source footprint of the tile (rows
// vstart[oy0] .. , byte columns of hstart[ox0] ..; replicate-clamped) is staged in shared
// memory, the vertical pass runs over it on 4-byte words with IDP.2A (two taps x
// 16-bit coefficients per instruction, int32 accumulate — exact), its u8 result stays in
// shared memory, and the horizontal pass reads that and writes the output tile once.
// The source is read from HBM/L2 once per tile; the intermediate never leaves the SM.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/irp_spec.h"

namespace irp {

constexpr int kResizeThreads = 256;
constexpr int kCoefStride = IRP_MAX_TAPS + 1;  // int16 per phase row (even count, so pairs load as u32)

struct AxisPlan {            // device pointers into the plan arena
  const int32_t* start;      // [out] first tap (may be < 0 / beyond the edge: clamped)
  const int32_t* phase;      // [out] 0..64
  const int16_t* coef;       // [65][kCoefStride]
  int n;                     // taps; 0 = identity on this axis
  int pad;
};

struct ResizeJob {
  const uint8_t* src;        // ORIENTED source (orientation already applied), interleaved
  unsigned long long src_pitch;
  uint8_t* dst;
  unsigned long long dst_pitch;
  int sw, sh;                // oriented source dims
  int c;                     // source channels (1, 3, 4)
  int dc;                    // destination channels (1 or 3)
  int dw, dh;                // resized dims
  int dst_x0, dst_y0;        // placement inside dst (fusion canvas), else 0
  int expand_grey;           // 1: C == 1 source replicated to 3 destination channels (fusion canvas)
  int tow, toh;              // output tile dims chosen by the host
  int tiles_x, tiles_y, tile_base;
  int src_rows_max, src_rowbytes_max;  // shared-memory tile bounds chosen by the host
  int aligned4;              // src base and pitch multiples of 4
  AxisPlan v, h;
};

__device__ __forceinline__ int dp2a_lo_s16_u8(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp2a_hi_s16_u8(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t fixed_round_u8(int v) {  // unsigned_fixed_round + clip
  v >>= IRP_INTERP_SHIFT;
  return (uint32_t)min(max(v, 0), 255);
}

__device__ __forceinline__ int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// gather-based orientation (P2); one thread per destination pixel
template <int C>
__global__ void orient_kernel(const uint8_t* __restrict__ src, unsigned long long spitch, int w, int h, int orientation,
                              uint8_t* __restrict__ dst, unsigned long long dpitch, int ow, int oh) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ow || y >= oh) return;
  int sx, sy;
  switch (orientation) {
    default: sx = x; sy = y; break;
    case 2: sx = w - 1 - x; sy = y; break;
    case 3: sx = w - 1 - x; sy = h - 1 - y; break;
    case 4: sx = x; sy = h - 1 - y; break;
    case 5: sx = y; sy = x; break;
    case 6: sx = y; sy = h - 1 - x; break;
    case 7: sx = w - 1 - y; sy = h - 1 - x; break;
    case 8: sx = w - 1 - y; sy = x; break;
  }
  const uint8_t* s = src + (size_t)sy * spitch + (size_t)sx * C;
  uint8_t* d = dst + (size_t)y * dpitch + (size_t)x * C;
#pragma unroll
  for (int ch = 0; ch < C; ch++) d[ch] = s[ch];
}

template <int C>
__global__ void __launch_bounds__(kResizeThreads)
resize_kernel(const ResizeJob* __restrict__ jobs, int n_jobs, int total_tiles) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t s_vcoef[IRP_MAX_TAPS / 2 + 1];
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    int ji = 0;
    while (ji + 1 < n_jobs && tile >= jobs[ji + 1].tile_base) ji++;
    const ResizeJob& J = jobs[ji];
    const int t = tile - J.tile_base;
    const int ty = t / J.tiles_x, tx = t - ty * J.tiles_x;
    const int ox0 = tx * J.tow, oy0 = ty * J.toh;
    const int ow = min(J.tow, J.dw - ox0), oh = min(J.toh, J.dh - oy0);
    const int vn = J.v.n, hn = J.h.n;
    // source footprint (oriented coordinates, unclamped)
    const int sy0 = vn ? J.v.start[oy0] : oy0;
    const int sy1 = vn ? J.v.start[oy0 + oh - 1] + vn : oy0 + oh;  // exclusive
    const int sx0 = hn ? J.h.start[ox0] : ox0;
    const int sx1 = hn ? J.h.start[ox0 + ow - 1] + hn : ox0 + ow;  // exclusive
    const int bx0 = floordiv(sx0 * C, 4) * 4;                       // byte origin, multiple of 4
    const int rowbytes = ((sx1 * C - bx0) + 3) & ~3;
    const int nrows = sy1 - sy0;
    uint8_t* src_t = smem;                                        // [nrows][rowbytes]
    uint8_t* mid_t = smem + (size_t)J.src_rows_max * J.src_rowbytes_max;  // [oh][rowbytes]
    const int rowwords = rowbytes >> 2;
    const int src_rowbytes_total = J.sw * C;

    // ---- stage A: source tile -> shared memory (replicate clamp) ----
    for (int i = threadIdx.x; i < nrows * rowwords; i += kResizeThreads) {
      const int r = i / rowwords, wj = i - r * rowwords;
      const int gy = min(max(sy0 + r, 0), J.sh - 1);
      const uint8_t* rp = J.src + (size_t)gy * J.src_pitch;
      const int b = bx0 + wj * 4;
      uint32_t word;
      if (J.aligned4 && b >= 0 && b + 4 <= src_rowbytes_total) {
        word = *reinterpret_cast<const uint32_t*>(rp + b);
      } else {
        word = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          int bb = b + k;
          int px = floordiv(bb, C), ch = bb - px * C;
          px = min(max(px, 0), J.sw - 1);
          word |= (uint32_t)rp[px * C + ch] << (8 * k);
        }
      }
      reinterpret_cast<uint32_t*>(src_t)[(size_t)r * rowwords + wj] = word;
    }
    __syncthreads();

    // ---- stage B: vertical reduce (reducev) on words, IDP.2A ----
    if (vn) {
      const int npairs = (vn + 1) >> 1;
      for (int orow = 0; orow < oh; orow++) {
        const int o = oy0 + orow;
        const int s = J.v.start[o] - sy0;
        const uint32_t* cp = reinterpret_cast<const uint32_t*>(J.v.coef + (size_t)J.v.phase[o] * kCoefStride);
        if (threadIdx.x < npairs) s_vcoef[threadIdx.x] = cp[threadIdx.x];
        __syncthreads();
        for (int wj = threadIdx.x; wj < rowwords; wj += kResizeThreads) {
          const uint32_t* colp = reinterpret_cast<const uint32_t*>(src_t) + (size_t)s * rowwords + wj;
          int a0 = 1 << (IRP_INTERP_SHIFT - 1), a1 = a0, a2 = a0, a3 = a0;
          for (int p = 0; p < npairs; p++) {
            uint32_t wa = colp[(size_t)(2 * p) * rowwords];
            // an odd tap count leaves the last pair's second coefficient 0: reading the row
            // below is harmless as long as it exists in the tile, so clamp the row index
            int r2 = min(s + 2 * p + 1, nrows - 1) - s;
            uint32_t wb = colp[(size_t)r2 * rowwords];
            uint32_t lo = __byte_perm(wa, wb, 0x5140), hi = __byte_perm(wa, wb, 0x7362);
            uint32_t c2 = s_vcoef[p];
            a0 = dp2a_lo_s16_u8(c2, lo, a0);
            a1 = dp2a_hi_s16_u8(c2, lo, a1);
            a2 = dp2a_lo_s16_u8(c2, hi, a2);
            a3 = dp2a_hi_s16_u8(c2, hi, a3);
          }
          uint32_t out = fixed_round_u8(a0) | (fixed_round_u8(a1) << 8) | (fixed_round_u8(a2) << 16) |
                         (fixed_round_u8(a3) << 24);
          reinterpret_cast<uint32_t*>(mid_t)[(size_t)orow * rowwords + wj] = out;
        }
        __syncthreads();
      }
    } else {
      for (int i = threadIdx.x; i < oh * rowwords; i += kResizeThreads)
        reinterpret_cast<uint32_t*>(mid_t)[i] = reinterpret_cast<const uint32_t*>(src_t)[i];
      __syncthreads();
    }

    // ---- stage C: horizontal reduce (reduceh) + normalise + store ----
    for (int i = threadIdx.x; i < oh * ow; i += kResizeThreads) {
      const int orow = i / ow, ocol = i - orow * ow;
      const int o = ox0 + ocol;
      const uint8_t* rp = mid_t + (size_t)orow * rowbytes;
      uint32_t v[C];
      if (hn) {
        const int16_t* cf = J.h.coef + (size_t)J.h.phase[o] * kCoefStride;
        const int base = J.h.start[o] * C - bx0;
        int acc[C];
#pragma unroll
        for (int ch = 0; ch < C; ch++) acc[ch] = 1 << (IRP_INTERP_SHIFT - 1);
        for (int k = 0; k < hn; k++) {
          const int cv = cf[k];
#pragma unroll
          for (int ch = 0; ch < C; ch++) acc[ch] += cv * (int)rp[base + k * C + ch];
        }
#pragma unroll
        for (int ch = 0; ch < C; ch++) v[ch] = fixed_round_u8(acc[ch]);
      } else {
        const int base = o * C - bx0;
#pragma unroll
        for (int ch = 0; ch < C; ch++) v[ch] = rp[base + ch];
      }
      uint8_t* d = J.dst + (size_t)(J.dst_y0 + oy0 + orow) * J.dst_pitch + (size_t)(J.dst_x0 + o) * J.dc;
      if (C == 4) {  // libvips flatten on black: p * a / 255, integer
        d[0] = (uint8_t)((v[0] * v[3]) / 255u);
        d[1] = (uint8_t)((v[1] * v[3]) / 255u);
        d[2] = (uint8_t)((v[2] * v[3]) / 255u);
      } else if (C == 1) {
        if (J.expand_grey) {
          d[0] = d[1] = d[2] = (uint8_t)v[0];
        } else {
          d[0] = (uint8_t)v[0];
        }
      } else {
#pragma unroll
        for (int ch = 0; ch < C; ch++) d[ch] = (uint8_t)v[ch];
      }
    }
    __syncthreads();
  }
}

}  // namespace irp
