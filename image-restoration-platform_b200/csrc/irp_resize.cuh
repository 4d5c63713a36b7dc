// irp_resize.cuh — preprocess pixel stages (sm_100a):
//   P2 EXIF auto-orient  (reference: server-node/src/middleware/imagePreprocess.js:42, sharp .rotate())
//   P3 lanczos3 fit-inside resize = libvips reducev then reduceh, 12-bit fixed point,
//      u8 between passes                                   (imagePreprocess.js:46-53)
//   P4 normalise to raw u8 (RGBA flattened on black)       (imagePreprocess.js:57-63, pixel part)
//   P5 placement on a 2048x2048 fusion canvas              
//
// One CTA produces one output tile (tow x toh pixels).  Data layout on the SM:
//   stage A  the tile's source footprint is pulled with 128-bit loads, de-interleaved to planes
//            (PRMT) and stored ROW-PAIR-INTERLEAVED: for source rows (2q, 2q+1) a shared-memory
//            word holds (a_k, b_k, a_k+1, b_k+1).  Replicate clamping happens here, once.
//   stage B  reducev: a thread owns 4 byte-columns of one plane; per tap PAIR it issues one LDS.64
//            and four IDP.2A (2 taps x s16 coefficient x u8 pixel, int32 accumulate — exact).  Odd
//            window starts are absorbed by shifting the coefficient pairs, not the data.
//            The u8 result stays in shared memory (planar).
//   stage C  reduceh: a thread owns one output column, keeps its (alignment-shifted) coefficient
//            pairs in registers for the whole tile, and walks rows x planes with LDS.32 + IDP.2A.
// The source is read from HBM/L2 once per tile and the intermediate never leaves the SM.
// No tensor cores: the fixed-point semantics are kept bit-exact on the integer pipes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/irp_spec.h"

namespace irp {

constexpr int kResizeThreads = 256;
constexpr int kCoefStride = IRP_MAX_TAPS + 1;  // int16 per phase row
constexpr int kSrcCols = 160;                  // source pixel columns a tile may touch
constexpr int kPairPitch = kSrcCols * 2;       // bytes per pair-row per plane
constexpr int kMidPitch = kSrcCols;            // bytes per row per plane of the reducev output
constexpr int kWordCols = kSrcCols / 4;        // 4-pixel word columns per plane
constexpr int kMaxPairs = (IRP_MAX_TAPS + 1) / 2;      // 13 coefficient pairs (leading zero included)
constexpr int kVtabStride = 16;                         // words per output row in the vertical table
constexpr int kMaxHPairs = (IRP_MAX_TAPS + 3 + 1) / 2;  // 14 pairs once shifted to word alignment
constexpr int kMaxTow = 64, kMaxToh = 32;

struct AxisPlan {            // device pointers into the plan arena
  const int32_t* start;      // [out] first tap (may be < 0 / beyond the edge: clamped)
  const int32_t* phase;      // [out] 0..64
  const int16_t* coef;       // [65][kCoefStride], zero padded
  // coefficient PAIRS (lo16 = tap 2p - shift, hi16 = tap 2p + 1 - shift), pre-shifted on the host:
  const uint32_t* vpairs;    // [65][2][16]  shift = window start parity (row pairs are even-aligned)
  const uint32_t* hpairs;    // [65][4][16]  shift = window start byte offset within its 32-bit word
  int n;                     // taps (1 = identity: coef 4096)
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
  int tow, toh;              // output tile dims chosen by the host (<= 64 x 32)
  int tiles_x, tiles_y, tile_base;
  int pairrows_max;          // shared-memory pair-rows reserved per plane
  int aligned16;             // src base and pitch are multiples of 16
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
// pack sat_u8(a) << 8 | sat_u8(b) into the low half, low half of c into the high half
__device__ __forceinline__ uint32_t pack_sat_u8(int a, int b, uint32_t c) {
  uint32_t d;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint4 ldg_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// 48 interleaved RGB bytes -> planar words R[4], G[4], B[4]
__device__ __forceinline__ void deinterleave48(const uint4& v0, const uint4& v1, const uint4& v2, uint32_t* R, uint32_t* G,
                                               uint32_t* B) {
  const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint32_t w0 = w[3 * k], w1 = w[3 * k + 1], w2 = w[3 * k + 2];
    R[k] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);
    G[k] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);
    B[k] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);
  }
}

// reducev for one 4-byte column of one plane over output rows [r_begin, r_end): straight-line code
// for NPV tap pairs.  vtab row = 13 coefficient pairs (+ pad) and, in word 15, the pair-row byte offset.
template <int NPV>
__device__ __forceinline__ void vpass_column(const uint8_t* sp, uint8_t* mp, const uint32_t* vtab, int r_begin, int r_end) {
  for (int r = r_begin; r < r_end; r++) {
    const uint32_t* vt = vtab + r * kVtabStride;
    uint32_t cp[16];
#pragma unroll
    for (int q4 = 0; q4 < (NPV + 3) / 4; q4++) {
      const uint4 c4 = reinterpret_cast<const uint4*>(vt)[q4];
      cp[4 * q4] = c4.x; cp[4 * q4 + 1] = c4.y; cp[4 * q4 + 2] = c4.z; cp[4 * q4 + 3] = c4.w;
    }
    const uint8_t* p = sp + vt[15];
    uint2 w[NPV];
#pragma unroll
    for (int pp = 0; pp < NPV; pp++) w[pp] = *reinterpret_cast<const uint2*>(p + pp * kPairPitch);
    int a0 = 1 << (IRP_INTERP_SHIFT - 1), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
    for (int pp = 0; pp < NPV; pp++) {
      a0 = dp2a_lo_s16_u8(cp[pp], w[pp].x, a0);
      a1 = dp2a_hi_s16_u8(cp[pp], w[pp].x, a1);
      a2 = dp2a_lo_s16_u8(cp[pp], w[pp].y, a2);
      a3 = dp2a_hi_s16_u8(cp[pp], w[pp].y, a3);
    }
    const uint32_t hi = pack_sat_u8(a3 >> IRP_INTERP_SHIFT, a2 >> IRP_INTERP_SHIFT, 0u);
    *reinterpret_cast<uint32_t*>(mp + r * kMidPitch) = pack_sat_u8(a1 >> IRP_INTERP_SHIFT, a0 >> IRP_INTERP_SHIFT, hi);
  }
}

// Two output rows per window load: rows r and r+1 start D pair-rows apart (D = 0, 1 or 2 for shrink < 4),
// their windows overlap in NPV - D pairs, so the union (NPV + D LDS.64) serves both: ~43 % fewer
// shared-memory reads than two single-row passes.
template <int NPV, int D>
__device__ __forceinline__ void vpass_two_rows(const uint8_t* sp, uint8_t* mp, const uint32_t* vtab, int r) {
  const uint32_t* vt0 = vtab + r * kVtabStride;
  const uint32_t* vt1 = vt0 + kVtabStride;
  uint32_t c0[16], c1[16];
#pragma unroll
  for (int q4 = 0; q4 < (NPV + 3) / 4; q4++) {
    const uint4 a = reinterpret_cast<const uint4*>(vt0)[q4], b = reinterpret_cast<const uint4*>(vt1)[q4];
    c0[4 * q4] = a.x; c0[4 * q4 + 1] = a.y; c0[4 * q4 + 2] = a.z; c0[4 * q4 + 3] = a.w;
    c1[4 * q4] = b.x; c1[4 * q4 + 1] = b.y; c1[4 * q4 + 2] = b.z; c1[4 * q4 + 3] = b.w;
  }
  const uint8_t* p = sp + vt0[15];
  uint2 w[NPV + D];
#pragma unroll
  for (int pp = 0; pp < NPV + D; pp++) w[pp] = *reinterpret_cast<const uint2*>(p + pp * kPairPitch);
  int a0 = 1 << (IRP_INTERP_SHIFT - 1), a1 = a0, a2 = a0, a3 = a0, b0 = a0, b1 = a0, b2 = a0, b3 = a0;
#pragma unroll
  for (int pp = 0; pp < NPV; pp++) {
    a0 = dp2a_lo_s16_u8(c0[pp], w[pp].x, a0);
    a1 = dp2a_hi_s16_u8(c0[pp], w[pp].x, a1);
    a2 = dp2a_lo_s16_u8(c0[pp], w[pp].y, a2);
    a3 = dp2a_hi_s16_u8(c0[pp], w[pp].y, a3);
    b0 = dp2a_lo_s16_u8(c1[pp], w[pp + D].x, b0);
    b1 = dp2a_hi_s16_u8(c1[pp], w[pp + D].x, b1);
    b2 = dp2a_lo_s16_u8(c1[pp], w[pp + D].y, b2);
    b3 = dp2a_hi_s16_u8(c1[pp], w[pp + D].y, b3);
  }
  const uint32_t ha = pack_sat_u8(a3 >> IRP_INTERP_SHIFT, a2 >> IRP_INTERP_SHIFT, 0u);
  *reinterpret_cast<uint32_t*>(mp + r * kMidPitch) = pack_sat_u8(a1 >> IRP_INTERP_SHIFT, a0 >> IRP_INTERP_SHIFT, ha);
  const uint32_t hb = pack_sat_u8(b3 >> IRP_INTERP_SHIFT, b2 >> IRP_INTERP_SHIFT, 0u);
  *reinterpret_cast<uint32_t*>(mp + (r + 1) * kMidPitch) = pack_sat_u8(b1 >> IRP_INTERP_SHIFT, b0 >> IRP_INTERP_SHIFT, hb);
}

// rows [r_begin, r_end) of one 4-byte column: pairs of rows where their windows are <= 2 pair-rows apart
template <int NPV>
__device__ __forceinline__ void vpass_rows(const uint8_t* sp, uint8_t* mp, const uint32_t* vtab, int r_begin, int r_end) {
  int r = r_begin;
  for (; r + 1 < r_end; r += 2) {
    const int d = (int)(vtab[(r + 1) * kVtabStride + 15] - vtab[r * kVtabStride + 15]) / kPairPitch;  // uniform
    if (d == 1)
      vpass_two_rows<NPV, 1>(sp, mp, vtab, r);
    else if (d == 0)
      vpass_two_rows<NPV, 0>(sp, mp, vtab, r);
    else if (d == 2)
      vpass_two_rows<NPV, 2>(sp, mp, vtab, r);
    else
      vpass_column<NPV>(sp, mp, vtab, r, r + 2);
  }
  if (r < r_end) vpass_column<NPV>(sp, mp, vtab, r, r_end);
}

// reduceh for one output column over rows [r_begin, r_end) and all planes; NWH aligned words per window.
template <int C, int NWH>
__device__ __forceinline__ void hpass_column(const uint8_t* mbase, int mid_plane, const uint32_t* hp, uint8_t* d,
                                             unsigned long long dst_pitch, int r_begin, int r_end, int expand_grey) {
  uint32_t cph[2 * NWH + 2];
#pragma unroll
  for (int q4 = 0; q4 < (2 * NWH + 3) / 4; q4++) {
    const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(hp) + q4);
    if (4 * q4 < 2 * NWH + 2) cph[4 * q4] = c4.x;
    if (4 * q4 + 1 < 2 * NWH + 2) cph[4 * q4 + 1] = c4.y;
    if (4 * q4 + 2 < 2 * NWH + 2) cph[4 * q4 + 2] = c4.z;
    if (4 * q4 + 3 < 2 * NWH + 2) cph[4 * q4 + 3] = c4.w;
  }
  for (int r = r_begin; r < r_end; r++) {
    uint32_t v[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) {
      const uint32_t* mp = reinterpret_cast<const uint32_t*>(mbase + ch * mid_plane + r * kMidPitch);
      uint32_t w[NWH];
#pragma unroll
      for (int wv = 0; wv < NWH; wv++) w[wv] = mp[wv];
      int acc = 1 << (IRP_INTERP_SHIFT - 1);
#pragma unroll
      for (int wv = 0; wv < NWH; wv++) {
        acc = dp2a_lo_s16_u8(cph[2 * wv], w[wv], acc);
        acc = dp2a_hi_s16_u8(cph[2 * wv + 1], w[wv], acc);
      }
      v[ch] = pack_sat_u8(0, acc >> IRP_INTERP_SHIFT, 0u);
    }
    uint8_t* dp = d + (size_t)r * dst_pitch;
    if (C == 4) {  // libvips flatten on black: p * a / 255, integer
      dp[0] = (uint8_t)((v[0] * v[3]) / 255u);
      dp[1] = (uint8_t)((v[1] * v[3]) / 255u);
      dp[2] = (uint8_t)((v[2] * v[3]) / 255u);
    } else if (C == 1) {
      if (expand_grey) {
        dp[0] = dp[1] = dp[2] = (uint8_t)v[0];
      } else {
        dp[0] = (uint8_t)v[0];
      }
    } else {
#pragma unroll
      for (int ch = 0; ch < C; ch++) dp[ch] = (uint8_t)v[ch];
    }
  }
}

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

// Identity geometry (source already <= 2048 px and 3 -> 3 or 1 -> 1 channels): the preprocess output
// is the oriented source, so the whole batch's identity jobs are one launch of row copies (one warp per
// row, the widest access both addresses allow).
struct CopyJob {
  const uint8_t* src;
  unsigned long long spitch;
  uint8_t* dst;
  unsigned long long dpitch;
  int row_bytes, rows, row_base, pad;
};

// vips_resize's integer pre-shrink (libvips resample/resize.c: shrinkv then shrinkh with "ceil"):
// kv rows then kh columns averaged, (sum + k / 2) / k each, u8 between the passes, edge replicated.  Only axes that
// shrink 4x or more come here (sides beyond 8192 px), so one thread per output byte is enough.
template <int C>
__global__ void __launch_bounds__(256) box_shrink_kernel(const uint8_t* __restrict__ src, size_t spitch, int w, int h, int kh, int kv,
                                                         uint8_t* __restrict__ dst, size_t dpitch, int ow, int oh) {
  const int xb = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (xb >= ow * C || y >= oh) return;
  const int x = xb / C, ch = xb - x * C;
  int acc = 0;
  for (int i = 0; i < kh; i++) {
    const int sx = min(x * kh + i, w - 1);
    int col = 0;
    for (int j = 0; j < kv; j++) col += __ldg(src + (size_t)min(y * kv + j, h - 1) * spitch + (size_t)sx * C + ch);
    acc += (col + kv / 2) / kv;
  }
  dst[(size_t)y * dpitch + xb] = (uint8_t)((acc + kh / 2) / kh);
}

__global__ void __launch_bounds__(256) copy_rows_kernel(const CopyJob* __restrict__ jobs, int n_jobs, int total_rows) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  int ji = 0;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < total_rows; row += nwarps) {
    while (ji + 1 < n_jobs && row >= __ldg(&jobs[ji + 1].row_base)) ji++;
    const CopyJob J = jobs[ji];
    const int r = row - J.row_base;
    const uint8_t* sp = J.src + (size_t)r * J.spitch;
    uint8_t* dp = J.dst + (size_t)r * J.dpitch;
    const uintptr_t both = (uintptr_t)sp | (uintptr_t)dp;
    int done = 0;
    if ((both & 15) == 0) {
      const int nv = J.row_bytes >> 4;
      for (int k = lane; k < nv; k += 32) reinterpret_cast<uint4*>(dp)[k] = __ldg(reinterpret_cast<const uint4*>(sp) + k);
      done = nv << 4;
    } else if ((both & 3) == 0) {
      const int nv = J.row_bytes >> 2;
      for (int k = lane; k < nv; k += 32) reinterpret_cast<uint32_t*>(dp)[k] = __ldg(reinterpret_cast<const uint32_t*>(sp) + k);
      done = nv << 2;
    }
    for (int k = done + lane; k < J.row_bytes; k += 32) dp[k] = __ldg(sp + k);
  }
}

// Orientation for 3-channel images through a shared-memory tile (the gather above reads one source
// ROW per lane for the transposing orientations 5-8: 32 sectors per load).  A CTA owns a 64x32-pixel
// DESTINATION tile: phase 1 pulls the matching source region (64x32, or 32x64 when transposed) with
// coalesced 32-bit loads into a tile whose odd word pitch makes column walks conflict-free; phase 2
// lets every thread build whole destination words (4 bytes = parts of 2 pixels) from byte reads of
// that tile and store them coalesced.  The destination scratch has a 16-byte pitch, so the word that
// straddles the right edge of the image spills into padding, never into the next row's pixels.
constexpr int kOrTw = 64, kOrTh = 32, kOrThreads = 192, kOrPitchW = 51;   // 51 words = 204 bytes per tile row

__global__ void __launch_bounds__(kOrThreads)
orient3_kernel(const uint8_t* __restrict__ src, unsigned long long spitch, int w, int h, int orientation, uint8_t* __restrict__ dst,
               unsigned long long dpitch, int ow, int oh) {
  __shared__ uint32_t tile[64 * kOrPitchW];
  const int x0 = blockIdx.x * kOrTw, y0 = blockIdx.y * kOrTh;
  const int tw = min(kOrTw, ow - x0), th = min(kOrTh, oh - y0);
  const bool transposed = orientation >= 5;
  // destination (x, y) -> source (sx, sy); flip_x / flip_y act on the SOURCE axes
  const bool flip_sx = orientation == 2 || orientation == 3 || orientation == 7 || orientation == 8;
  const bool flip_sy = orientation == 3 || orientation == 4 || orientation == 6 || orientation == 7;
  // source region of this tile: columns [sx0, sx0 + ncx), rows [sy0, sy0 + ncy)
  const int ncx = transposed ? th : tw, ncy = transposed ? tw : th;
  const int dx0 = transposed ? y0 : x0, dy0 = transposed ? x0 : y0;       // tile origin along source x / y before flips
  const int sx0 = flip_sx ? w - dx0 - ncx : dx0, sy0 = flip_sy ? h - dy0 - ncy : dy0;
  // phase 1: rows of the region as words, from the word that holds its first byte
  const int b0 = sx0 * 3, wb0 = b0 & ~3, nwords = ((b0 + ncx * 3 + 3) >> 2) - (wb0 >> 2);
  const int boff = b0 - wb0;
  for (int e = threadIdx.x; e < ncy * nwords; e += kOrThreads) {
    const int r = e / nwords, k = e - r * nwords;
    tile[r * kOrPitchW + k] = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)(sy0 + r) * spitch + wb0) + k);
  }
  __syncthreads();
  // phase 2: thread = one word column of the destination tile (48 per row) x one of 4 row groups
  const int wq = threadIdx.x % 48, rg = threadIdx.x / 48;
  if (wq * 4 >= tw * 3) return;
  const uint8_t* tb = reinterpret_cast<const uint8_t*>(tile);
  int off[4], step;   // byte offset inside the tile of the word's four bytes for destination row 0, and the per-row step
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const int B = wq * 4 + b;
    const int p = min(B / 3, tw - 1), ch = B - (B / 3) * 3;               // destination pixel (clamped into the tile) and channel
    if (!transposed) {
      const int cx = flip_sx ? ncx - 1 - p : p;
      off[b] = boff + cx * 3 + ch + (flip_sy ? (ncy - 1) * kOrPitchW * 4 : 0);
    } else {                                                              // destination x walks source rows
      const int cy = flip_sy ? ncy - 1 - p : p;
      off[b] = boff + cy * kOrPitchW * 4 + ch + (flip_sx ? (ncx - 1) * 3 : 0);
    }
  }
  step = !transposed ? (flip_sy ? -kOrPitchW * 4 : kOrPitchW * 4) : (flip_sx ? -3 : 3);   // destination y walks source rows / columns
  uint8_t* drow = dst + (size_t)y0 * dpitch + (size_t)x0 * 3 + wq * 4;
  for (int ty = rg; ty < th; ty += 4) {
    const int o = ty * step;
    const uint32_t v = tb[off[0] + o] | (tb[off[1] + o] << 8) | (tb[off[2] + o] << 16) | (tb[off[3] + o] << 24);
    *reinterpret_cast<uint32_t*>(drow + (size_t)ty * dpitch) = v;
  }
}

template <int C>
__global__ void __launch_bounds__(kResizeThreads)
resize_kernel(const ResizeJob* __restrict__ jobs, int n_jobs, int total_tiles) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ __align__(16) uint32_t s_vtab[kMaxToh * kVtabStride];
  __shared__ __align__(16) ResizeJob s_job;   // the current job, so its fields are LDS not LDG
  const int tid = threadIdx.x;
  int ji = 0, loaded = -1;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    while (ji + 1 < n_jobs && tile >= jobs[ji + 1].tile_base) ji++;
    if (ji != loaded) {  // uniform: every thread sees the same tile sequence
      __syncthreads();
      const uint32_t* src32 = reinterpret_cast<const uint32_t*>(jobs + ji);
      for (int k = tid; k < (int)(sizeof(ResizeJob) / 4); k += kResizeThreads) reinterpret_cast<uint32_t*>(&s_job)[k] = src32[k];
      loaded = ji;
      __syncthreads();
    }
    const ResizeJob& J = s_job;
    const int t = tile - J.tile_base;
    const int ty = t / J.tiles_x, tx = t - ty * J.tiles_x;
    const int ox0 = tx * J.tow, oy0 = ty * J.toh;
    const int ow = min(J.tow, J.dw - ox0), oh = min(J.toh, J.dh - oy0);
    const int vn = J.v.n, hn = J.h.n;
    const int sy0 = J.v.start[oy0], sy1 = J.v.start[oy0 + oh - 1] + vn;  // [sy0, sy1) source rows, unclamped
    const int sx0 = J.h.start[ox0], sx1 = J.h.start[ox0 + ow - 1] + hn;  // [sx0, sx1) source columns, unclamped
    const int sy0e = sy0 & ~1;                                                                            // pair-row origin (even)
    const int npairrows = (sy1 - sy0e + 1) >> 1;
    const bool fast_img = (C == 3) && J.aligned16;
    const int sx0a = fast_img ? (sx0 & ~15) : sx0;                         // column origin
    const int ncols = sx1 - sx0a;                                         // <= kSrcCols (host guarantees)
    const int src_plane = J.pairrows_max * kPairPitch;
    const int mid_plane = kMaxToh * kMidPitch;
    uint8_t* src_t = smem;                      // [C][pairrows_max][kPairPitch]
    uint8_t* mid_t = smem + (size_t)C * src_plane;  // [C][kMaxToh][kMidPitch]

    // ---- vertical table: per output row, 13 coefficient pairs (even-row aligned) + pair-row offset ----
    for (int e = tid; e < oh * kVtabStride; e += kResizeThreads) {
      const int r = e >> 4, k = e & 15;
      const int o = oy0 + r;
      const int s0 = J.v.start[o];
      const int se = s0 & ~1, lead = s0 & 1;
      s_vtab[e] = k == 15 ? (uint32_t)(((se - sy0e) >> 1) * kPairPitch)
                          : J.v.vpairs[((size_t)J.v.phase[o] * 2 + lead) * 16 + k];
    }

    // ---- stage A: source footprint -> planar, row-pair-interleaved shared tile ----
    {
      const int nchunks = (ncols + 15) >> 4;
      const int nitems = npairrows * nchunks;
      const uint32_t inv_chunks = 65536u / (uint32_t)nchunks + 1u;  // exact for item < 65536 / nchunks... (items <= ~1300)
#ifndef IRP_SKIP_A
      for (int item = tid; item < nitems; item += kResizeThreads) {
        const int q = (int)(((uint32_t)item * inv_chunks) >> 16), k = item - q * nchunks;
        const int ra = min(max(sy0e + 2 * q, 0), J.sh - 1), rb = min(max(sy0e + 2 * q + 1, 0), J.sh - 1);
        const uint8_t* pa = J.src + (size_t)ra * J.src_pitch;
        const uint8_t* pb = J.src + (size_t)rb * J.src_pitch;
        const int x = sx0a + 16 * k;
        uint8_t* dst = src_t + q * kPairPitch + 32 * k;
        bool done = false;
        if constexpr (C == 3) {
          if (fast_img && x >= 0 && x + 16 <= J.sw) {
            uint32_t Ra[4], Ga[4], Ba[4], Rb[4], Gb[4], Bb[4];
            {
              const uint4 a0 = ldg_v4(pa + x * 3), a1 = ldg_v4(pa + x * 3 + 16), a2 = ldg_v4(pa + x * 3 + 32);
              const uint4 b0 = ldg_v4(pb + x * 3), b1 = ldg_v4(pb + x * 3 + 16), b2 = ldg_v4(pb + x * 3 + 32);
              deinterleave48(a0, a1, a2, Ra, Ga, Ba);
              deinterleave48(b0, b1, b2, Rb, Gb, Bb);
            }
#pragma unroll
            for (int h2 = 0; h2 < 2; h2++) {
              const int k0 = 2 * h2, k1 = k0 + 1;
              *reinterpret_cast<uint4*>(dst + 16 * h2) =
                  make_uint4(__byte_perm(Ra[k0], Rb[k0], 0x5140), __byte_perm(Ra[k0], Rb[k0], 0x7362),
                             __byte_perm(Ra[k1], Rb[k1], 0x5140), __byte_perm(Ra[k1], Rb[k1], 0x7362));
              *reinterpret_cast<uint4*>(dst + src_plane + 16 * h2) =
                  make_uint4(__byte_perm(Ga[k0], Gb[k0], 0x5140), __byte_perm(Ga[k0], Gb[k0], 0x7362),
                             __byte_perm(Ga[k1], Gb[k1], 0x5140), __byte_perm(Ga[k1], Gb[k1], 0x7362));
              *reinterpret_cast<uint4*>(dst + 2 * src_plane + 16 * h2) =
                  make_uint4(__byte_perm(Ba[k0], Bb[k0], 0x5140), __byte_perm(Ba[k0], Bb[k0], 0x7362),
                             __byte_perm(Ba[k1], Bb[k1], 0x5140), __byte_perm(Ba[k1], Bb[k1], 0x7362));
            }
            done = true;
          }
        }
        if (!done) {
          const int lim = min(16, ncols - 16 * k);
          for (int i = 0; i < lim; i++) {
            const int xc = min(max(x + i, 0), J.sw - 1);
#pragma unroll
            for (int ch = 0; ch < C; ch++)
              *reinterpret_cast<uint16_t*>(dst + ch * src_plane + 2 * i) =
                  (uint16_t)(pa[(size_t)xc * C + ch] | (pb[(size_t)xc * C + ch] << 8));
          }
        }
      }
#endif
    }
    __syncthreads();

    // ---- L2 prefetch of this CTA's next tile (same job only), so its stage A finds the lines in L2 ----
    {
      const int tn = t + (int)gridDim.x;
      if (tn < J.tiles_x * J.tiles_y) {
        const int tyn = tn / J.tiles_x, txn = tn - tyn * J.tiles_x;
        const int oxn = txn * J.tow, oyn = tyn * J.toh;
        const int ohn = min(J.toh, J.dh - oyn), own = min(J.tow, J.dw - oxn);
        const int ya = J.v.start[oyn], yb = J.v.start[oyn + ohn - 1] + vn;
        const int xa = max(J.h.start[oxn], 0) * C, xb = min(J.h.start[oxn + own - 1] + hn, J.sw) * C;
        const int lines = ((xb - (xa & ~127)) + 127) >> 7;   // 128-byte lines per row
        const int total = (yb - ya) * lines;
        for (int e = tid; e < total; e += kResizeThreads) {
          const int r = e / lines, l = e - r * lines;
          const int gy = min(max(ya + r, 0), J.sh - 1);
          const uint8_t* pp = J.src + (size_t)gy * J.src_pitch + (xa & ~127) + l * 128;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
        }
      }
    }

    // ---- stage B: reducev (vertical), LDS.64 + 4 x IDP.2A per tap pair ----
    {
      const int rg = tid >> 7;                   // two row groups
      const int rows_per = (oh + 1) >> 1;
      const int r_begin = rg * rows_per, r_end = min(oh, r_begin + rows_per);
      const int npv = (vn + 2) >> 1;             // pairs incl. a possible leading zero (uniform)
      const int nwc = (ncols + 3) >> 2;          // word columns in use
#ifndef IRP_SKIP_B
      for (int col = tid & 127; col < C * kWordCols; col += 128) {
        const int plane = col / kWordCols, j = col - plane * kWordCols;
        if (j >= nwc) continue;
        const uint8_t* sp = src_t + plane * src_plane + j * 8;
        uint8_t* mp = mid_t + plane * mid_plane + j * 4;
        switch (npv) {
          case 1: vpass_column<1>(sp, mp, s_vtab, r_begin, r_end); break;
          case 2: vpass_column<2>(sp, mp, s_vtab, r_begin, r_end); break;
          case 3: vpass_column<3>(sp, mp, s_vtab, r_begin, r_end); break;
          case 4: vpass_rows<4>(sp, mp, s_vtab, r_begin, r_end); break;
          case 5: vpass_rows<5>(sp, mp, s_vtab, r_begin, r_end); break;
          case 6: vpass_rows<6>(sp, mp, s_vtab, r_begin, r_end); break;
          case 7: vpass_rows<7>(sp, mp, s_vtab, r_begin, r_end); break;
          case 8: vpass_rows<8>(sp, mp, s_vtab, r_begin, r_end); break;
          case 9: vpass_rows<9>(sp, mp, s_vtab, r_begin, r_end); break;
          case 10: vpass_rows<10>(sp, mp, s_vtab, r_begin, r_end); break;
          case 11: vpass_column<11>(sp, mp, s_vtab, r_begin, r_end); break;
          case 12: vpass_column<12>(sp, mp, s_vtab, r_begin, r_end); break;
          default: vpass_column<13>(sp, mp, s_vtab, r_begin, r_end); break;
        }
      }
#endif
    }
    __syncthreads();

    // ---- stage C: reduceh (horizontal) + normalise + store ----
    {
      const int xcol = tid & 63, rgh = tid >> 6;  // four row groups
#ifndef IRP_SKIP_C
      if (xcol < ow) {
        const int o = ox0 + xcol;
        const int start = J.h.start[o] - sx0a;    // >= 0
        const uint32_t* hp = J.h.hpairs + ((size_t)J.h.phase[o] * 4 + (start & 3)) * 16;
        const int nwh = (hn + 3 + 3) >> 2;        // words per window once shifted to word alignment (uniform)
        const int rows_per = (oh + 3) >> 2;
        const int r_begin = rgh * rows_per, r_end = min(oh, r_begin + rows_per);
        const uint8_t* mbase = mid_t + (start >> 2) * 4;
        uint8_t* d = J.dst + (size_t)(J.dst_y0 + oy0) * J.dst_pitch + (size_t)(J.dst_x0 + o) * J.dc;
        const int eg = J.expand_grey;
        switch (nwh) {
          case 1: hpass_column<C, 1>(mbase, mid_plane, hp, d, J.dst_pitch, r_begin, r_end, eg); break;
          case 2: hpass_column<C, 2>(mbase, mid_plane, hp, d, J.dst_pitch, r_begin, r_end, eg); break;
          case 3: hpass_column<C, 3>(mbase, mid_plane, hp, d, J.dst_pitch, r_begin, r_end, eg); break;
          case 4: hpass_column<C, 4>(mbase, mid_plane, hp, d, J.dst_pitch, r_begin, r_end, eg); break;
          case 5: hpass_column<C, 5>(mbase, mid_plane, hp, d, J.dst_pitch, r_begin, r_end, eg); break;
          case 6: hpass_column<C, 6>(mbase, mid_plane, hp, d, J.dst_pitch, r_begin, r_end, eg); break;
          default: hpass_column<C, 7>(mbase, mid_plane, hp, d, J.dst_pitch, r_begin, r_end, eg); break;
        }
      }
#endif
    }
    __syncthreads();
  }
}

}  // namespace irp
