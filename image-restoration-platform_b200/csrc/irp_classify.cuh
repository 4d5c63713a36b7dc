// irp_classify.cuh — fused degradation-statistics kernel (sm_100a).
//
// One pass over the source pixels computes everything ClassifierService.analyze
// (reference: server-node/src/services/classifier.js:40-99) derives from them:
//   S0  per-channel sum / sum of squares            (classifier.js:52, sharp.stats)
//   G1  libvips B_W grey, integer restatement       (classifier.js:108,136,200)
//   A1  sum / sumsq of clip(Lap8 * grey)            (classifier.js:107-118)
//   A2  sum / sumsq of clip(Sharp9 * grey)          (classifier.js:135-145)
//   A4  pooled sum / sumsq of gaussblur(1) bytes    (classifier.js:296-300)
//   A5  stride-4 grid counts on clip(Lap4 * grey)   (classifier.js:310-337)
//   + luma histogram and 8x8 boundary step counts (additive diagnostics).
//
// Layout: a persistent grid walks 256x32-pixel tiles of the whole batch.  Per
// tile, stage 1 pulls 48-byte RGB segments with 128-bit loads, regroups them to
// planar R/G/B words (PRMT), accumulates channel moments with IDP.4A, maps the
// pixels to grey through shared-memory integer LUTs and writes planar + grey
// tiles (1-pixel replicate halo) to shared memory.  Stage 2 slides a 3-row
// window down 4-pixel strips of the grey tile in packed 16x2 SIMD; stage 3 does
// the separable 12/20/12 blur the same way (IDP.4A + IMAD.HI exact /11).
// Per-thread u32 accumulators live in registers across the tiles of one image
// and are flushed with one block reduction + 64-bit global atomics per image.
// HBM-bound byte work: no tensor cores by design (BASELINE.json north_star).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/irp_spec.h"

namespace irp {

constexpr int kTileW = 128;            // owned pixels per tile row
constexpr int kTileH = 32;             // owned rows per tile
constexpr int kRows = kTileH + 2;      // + replicate halo rows
constexpr int kPlanePitch = 16 + kTileW + 16;  // bytes: [..15 = left halo][256 px][0 = right halo ..]
constexpr int kGroupThreads = 128;       // one tile is worked on by a group of 4 warps ...
constexpr int kGroups = 2;               // ... and a CTA runs two groups on their own tiles, sharing the lookup tables
constexpr int kClassifyThreads = kGroupThreads * kGroups;
constexpr int kStrips = kTileW / 4;      // 4-pixel strips per tile row == threads per row group
constexpr int kSegPx = 16;             // pixels per stage-1 work item
constexpr int kSegsPerRow = kTileW / kSegPx;
constexpr int kFlushTiles = 16;        // keeps every per-WARP u32 sum below 2^32 (blur sumsq: 16*6.3e6*32)

// global accumulator slots (u64) per image
enum {
  ACC_SUM = 0,     // 4
  ACC_SUMSQ = 4,   // 4
  ACC_E1S = 8, ACC_E1Q = 9, ACC_E2S = 10, ACC_E2Q = 11,
  ACC_BS = 12, ACC_BQ = 13,
  ACC_SV = 14, ACC_SH = 15, ACC_BE0 = 16, ACC_BE1 = 17,
  ACC_COUNT = 18
};

struct ImgDev {
  const uint8_t* px;
  unsigned long long pitch;
  int w, h, c;
  int tiles_x, tiles_y;
  int tile_base;    // first global tile index of this image
  int aligned16;    // base and pitch are multiples of 16
  int slot;         // index of this image's accumulators / result
};

struct ClassifyTables {   // device copies of grey_tables.inc for the chosen luma mode
  uint32_t lut[3][256];
  uint32_t inv[4096];
};

// ---------------------------------------------------------------------------
// Shared-memory layout.  Everything is dynamic shared memory, carved at run time so that the
// randomly indexed tables sit on power-of-two boundaries of the .shared address space: a lookup is
// then   SHF (byte -> offset) ; LOP3 ((x & mask) | table_base) ; LDS   — no base add.
//   group-0 tiles | group-1 tiles | 16 KB-aligned: inv | 1 KB-aligned: lutR, lutG, lutB, hist x2 | red x2
// ---------------------------------------------------------------------------
constexpr int kTileBytes = kRows * kPlanePitch;
constexpr int kRedBytes = (kGroupThreads / 32) * ACC_COUNT * 4;

struct SmemMap {           // .shared-space byte addresses
  uint32_t tiles[kGroups]; // grey + C planes, contiguous, per group
  uint32_t inv, lut_r, lut_g, lut_b, hist[kGroups], red[kGroups], end;
};
__host__ __device__ inline SmemMap make_smem_map(uint32_t base, int C) {
  SmemMap m;
  const uint32_t tile_set = (uint32_t)(C + 1) * kTileBytes;
  // both tile sets first (they fit below the first 16 KB boundary for C <= 3 only partly; the map is generic)
  m.tiles[0] = base;
  m.tiles[1] = (base + tile_set + 15u) & ~15u;
  m.inv = (m.tiles[1] + tile_set + 16383u) & ~16383u;
  m.lut_r = m.inv + 16384u;
  m.lut_g = m.lut_r + 1024u;
  m.lut_b = m.lut_r + 2048u;
  m.hist[0] = m.lut_r + 3072u;
  m.hist[1] = m.lut_r + 4096u;
  m.red[0] = m.lut_r + 5120u;
  m.red[1] = m.red[0] + kRedBytes;
  m.end = m.red[1] + kRedBytes;
  return m;
}

template <int C>
struct Tiles {             // generic pointers into the same dynamic shared memory
  uint8_t* grey;           // [kRows][kPlanePitch]
  uint8_t* plane[C];       // [kRows][kPlanePitch] each
  uint32_t* hist;          // [256]
  uint32_t* red;           // [warps][ACC_COUNT]
  uint32_t a_lut_r, a_lut_g, a_lut_b, a_inv, a_hist;  // aligned .shared addresses of the tables
};

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void red_inc_shared(uint32_t addr) {
  asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}

// dp2a with unsigned 16-bit lanes in a and SIGNED bytes in b (low half of b)
__device__ __forceinline__ int dp2a_lo_u16_s8(uint32_t a, uint32_t b, int c) {
  int d;
  asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// byte j of a packed word -> byte offset into a 256 x u32 table
template <int J>
__device__ __forceinline__ uint32_t byte_off4(uint32_t w) {
  return J == 0 ? ((w << 2) & 0x3FCu) : ((w >> (8 * J - 2)) & 0x3FCu);
}
// integer restatement of libvips colourspace(B_W): returns inv[I >> 20] + (I & 0xFFFFF); grey = top byte.
// I = (I >> 20 << 20) + (I & 0xFFFFF), so with inv'[k] = inv[k] - (k << 20) the same value is inv'[I >> 20] + I.
template <int C>
__device__ __forceinline__ uint32_t grey_top_off(const Tiles<C>& T, uint32_t ro, uint32_t go, uint32_t bo) {
  const uint32_t I = lds_u32(ro | T.a_lut_r) + lds_u32(go | T.a_lut_g) + lds_u32(bo | T.a_lut_b);
  return lds_u32(((I >> 18) & 0x3FFCu) | T.a_inv) + I;   // the shared copy of inv[k] has k << 20 subtracted: no masking of I
}
template <int C>
__device__ __forceinline__ void hist_add_top(const Tiles<C>& T, uint32_t t) {
  red_inc_shared(((t >> 22) & 0x3FCu) | T.a_hist);
}


// barrier over the 128 threads of one group (named barriers 1..; barrier 0 stays __syncthreads)
__device__ __forceinline__ void group_barrier(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(kGroupThreads) : "memory");
}

template <int C>
struct Acc {
  uint32_t s[C], q[C];
  uint32_t e1s, e1q, e2s, e2q, bs, bq, sv, sh, be0, be1;
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < C; i++) s[i] = q[i] = 0;
    e1s = e1q = e2s = e2q = bs = bq = sv = sh = be0 = be1 = 0;
  }
};

template <int C>
__device__ void flush_acc(const Tiles<C>& T, Acc<C>& a, unsigned long long* gacc, uint32_t* ghist, int tid, int group) {
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t v[ACC_COUNT];
#pragma unroll
  for (int i = 0; i < ACC_COUNT; i++) v[i] = 0;
#pragma unroll
  for (int i = 0; i < C; i++) {
    v[ACC_SUM + i] = a.s[i];
    v[ACC_SUMSQ + i] = a.q[i];
  }
  v[ACC_E1S] = a.e1s; v[ACC_E1Q] = a.e1q; v[ACC_E2S] = a.e2s; v[ACC_E2Q] = a.e2q;
  v[ACC_BS] = a.bs; v[ACC_BQ] = a.bq; v[ACC_SV] = a.sv; v[ACC_SH] = a.sh;
  v[ACC_BE0] = a.be0; v[ACC_BE1] = a.be1;
#pragma unroll
  for (int i = 0; i < ACC_COUNT; i++) {
    const uint32_t r = __reduce_add_sync(0xffffffffu, v[i]);  // REDUX.SUM: one instruction per value
    if (lane == 0) T.red[warp * ACC_COUNT + i] = r;
  }
  group_barrier(group);
  if (tid < ACC_COUNT) {
    unsigned long long t = 0;
#pragma unroll
    for (int w = 0; w < kGroupThreads / 32; w++) t += T.red[w * ACC_COUNT + tid];
    if (t) atomicAdd(&gacc[tid], t);
  }
  for (int b = tid; b < 256; b += kGroupThreads) {
    const uint32_t hv = T.hist[b];
    if (hv) atomicAdd(&ghist[b], hv);
    T.hist[b] = 0;
  }
  a.clear();
  group_barrier(group);
}

// ---------------------------------------------------------------------------
// stage 1, generic item: any channel count, any alignment, image edges.
// One item = up to 16 pixels of one tile row.
// ---------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void stage1_slow(const Tiles<C>& T, Acc<C>& a, const ImgDev& im, int gy, int xbeg, int row, int col0,
                                            int npx, bool counted) {
  const uint8_t* rp = im.px + (size_t)gy * im.pitch;
  const int o = row * kPlanePitch + col0;
  for (int i = 0; i < npx; i++) {
    const int x = xbeg + i;
    const bool inside = x >= 0 && x < im.w;
    const int xc = min(max(x, 0), im.w - 1);
    uint32_t v[C];
#pragma unroll
    for (int ch = 0; ch < C; ch++) v[ch] = rp[(size_t)xc * C + ch];
    uint32_t g;
    if constexpr (C >= 3)
      g = grey_top_off(T, v[0] << 2, v[1] << 2, v[2] << 2) >> 24;
    else
      g = v[0];
#pragma unroll
    for (int ch = 0; ch < C; ch++) T.plane[ch][o + i] = (uint8_t)v[ch];
    T.grey[o + i] = (uint8_t)g;
    if (counted && inside) {
#pragma unroll
      for (int ch = 0; ch < C; ch++) {
        a.s[ch] += v[ch];
        a.q[ch] += v[ch] * v[ch];
      }
      red_inc_shared((g << 2) | T.a_hist);
    }
  }
}

// stage 1, fast item: C == 3, 48 aligned bytes fully inside the image.  COUNTED = the row belongs to
// the tile (moments + histogram); halo rows only feed the stencils.
template <bool COUNTED>
__device__ __forceinline__ void stage1_fast(const Tiles<3>& T, Acc<3>& a, const uint8_t* p, int row, int col0) {
  const uint4 v0 = ldg_nc_v4(p), v1 = ldg_nc_v4(p + 16), v2 = ldg_nc_v4(p + 32);
  const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
  uint32_t R[4], G[4], B[4], Y[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    // w0 = R0 G0 B0 R1 | w1 = G1 B1 R2 G2 | w2 = B2 R3 G3 B3
    const uint32_t w0 = w[3 * k], w1 = w[3 * k + 1], w2 = w[3 * k + 2];
    R[k] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);  // R0 R1 R2 | R3
    G[k] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);  // G0 G1 G2 | G3
    B[k] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);  // B0 B1 | B2 B3
  }
  if (COUNTED) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      a.s[0] = __dp4a(R[k], 0x01010101u, a.s[0]);
      a.s[1] = __dp4a(G[k], 0x01010101u, a.s[1]);
      a.s[2] = __dp4a(B[k], 0x01010101u, a.s[2]);
      a.q[0] = __dp4a(R[k], R[k], a.q[0]);
      a.q[1] = __dp4a(G[k], G[k], a.q[1]);
      a.q[2] = __dp4a(B[k], B[k], a.q[2]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint32_t t[4];
    t[0] = grey_top_off(T, byte_off4<0>(R[k]), byte_off4<0>(G[k]), byte_off4<0>(B[k]));
    t[1] = grey_top_off(T, byte_off4<1>(R[k]), byte_off4<1>(G[k]), byte_off4<1>(B[k]));
    t[2] = grey_top_off(T, byte_off4<2>(R[k]), byte_off4<2>(G[k]), byte_off4<2>(B[k]));
    t[3] = grey_top_off(T, byte_off4<3>(R[k]), byte_off4<3>(G[k]), byte_off4<3>(B[k]));
    if (COUNTED) {
#pragma unroll
      for (int j = 0; j < 4; j++) hist_add_top(T, t[j]);
    }
    Y[k] = __byte_perm(__byte_perm(t[0], t[1], 0x0073), __byte_perm(t[2], t[3], 0x0073), 0x5410);
  }
  const int o = row * kPlanePitch + col0;
  *reinterpret_cast<uint4*>(T.plane[0] + o) = make_uint4(R[0], R[1], R[2], R[3]);
  *reinterpret_cast<uint4*>(T.plane[1] + o) = make_uint4(G[0], G[1], G[2], G[3]);
  *reinterpret_cast<uint4*>(T.plane[2] + o) = make_uint4(B[0], B[1], B[2], B[3]);
  *reinterpret_cast<uint4*>(T.grey + o) = make_uint4(Y[0], Y[1], Y[2], Y[3]);
}

// ---------------------------------------------------------------------------
// stage 2: 3x3 stencils on the grey tile, packed 16x2.
// Thread = 4-pixel strip x 8 rows.  Pairs: A = (px0, px1), B = (px2, px3).
// ---------------------------------------------------------------------------
struct GreyRow {
  uint32_t cA, cB;    // centre pairs
  uint32_t hA, hB;    // horizontal 3-sums
  uint32_t p34;       // (px3, px4) for the column-boundary diagnostic
  uint32_t word;      // the 4 centre bytes
};

// Row layout: bytes 0..15 pad (15 = pixel x0-1), 16..271 = pixels 0..255, 272 = pixel x0+256.
// Strip s owns word 4+s; word 3+s ends with its left neighbour, word 5+s starts with its right one.
__device__ __forceinline__ GreyRow load_grey_row(const uint8_t* grey_row, int strip) {
  const uint32_t* rp = reinterpret_cast<const uint32_t*>(grey_row) + 3 + strip;
  const uint32_t l = rp[0], c = rp[1], r = rp[2];
  const uint32_t w0 = __funnelshift_r(l, c, 24);  // (l3, c0, c1, c2)
  const uint32_t w3 = __funnelshift_r(c, r, 16);  // (c2, c3, r0, r1)
  // 16x2 pairs (low lane = left pixel); selector 4 picks a zero byte from the second operand
  const uint32_t P_m1_0 = __byte_perm(w0, 0u, 0x4140), P_0_1 = __byte_perm(w0, 0u, 0x4241);
  const uint32_t P_1_2 = __byte_perm(w0, 0u, 0x4342), P_2_3 = __byte_perm(w3, 0u, 0x4140);
  const uint32_t P_3_4 = __byte_perm(w3, 0u, 0x4241);
  GreyRow g;
  g.word = c;
  g.cA = P_0_1;
  g.cB = P_2_3;
  g.hA = P_m1_0 + P_0_1 + P_1_2;
  g.hB = P_1_2 + P_2_3 + P_3_4;
  g.p34 = P_3_4;
  return g;
}

// clip(k*c - nb, 0, 255) per 16-bit lane, all operands non-negative lanes
__device__ __forceinline__ uint32_t clip_diff(uint32_t kc, uint32_t nb) {
  uint32_t mx = __vmaxu2(kc, nb);
  return __vminu2(mx - nb, 0x00FF00FFu);
}

// FULL = the tile lies strictly inside the image (x0 + 256 < W and y0 + 32 < H): no lane masks,
// no row / column validity tests.  Edge tiles take the masked instantiation.
template <int C, bool FULL, int PITCH = kPlanePitch>
__device__ __forceinline__ void stage2(const Tiles<C>& T, Acc<C>& a, int tid, int x0, int y0, int W, int H) {
  const int strip = tid & (kStrips - 1), rg = tid / kStrips;
  const int x = x0 + strip * 4;
  const int r0 = rg * 8;
  int nvalid = 4, nrows = 8;
  if (!FULL) {
    if (x >= W || y0 + r0 >= H) return;
    nvalid = min(4, W - x);
    nrows = min(8, H - (y0 + r0));
  }
  // byte masks for partially valid strips (pairs hold values in bytes 0 and 2)
  const uint32_t mA = nvalid >= 2 ? 0x00FF00FFu : 0x000000FFu;
  const uint32_t mB = nvalid >= 4 ? 0x00FF00FFu : (nvalid == 3 ? 0x000000FFu : 0u);
  GreyRow up = load_grey_row(T.grey + r0 * PITCH, strip), cur = load_grey_row(T.grey + (r0 + 1) * PITCH, strip);
  uint32_t prev_t0 = 0;  // E3 > 200 at (grid row, px0), consumed by the row below
  const bool col_boundary = (strip & 1) && (FULL || x + 4 < W);  // px3 | px4 straddle an 8-px column boundary
#pragma unroll 4
  for (int i = 0; i < 8; i++) {
    if (!FULL && i >= nrows) break;
    const GreyRow dn = load_grey_row(T.grey + (r0 + i + 2) * PITCH, strip);
    const uint32_t boxA = up.hA + cur.hA + dn.hA, boxB = up.hB + cur.hB + dn.hB;
    const uint32_t nineA = cur.cA * 9u, nineB = cur.cB * 9u;
    uint32_t e1A = clip_diff(nineA, boxA), e1B = clip_diff(nineB, boxB);
    uint32_t e2A = clip_diff(nineA + cur.cA, boxA), e2B = clip_diff(nineB + cur.cB, boxB);
    if (!FULL) {
      e1A &= mA; e1B &= mB; e2A &= mA; e2B &= mB;
    }
    const uint32_t e1 = __byte_perm(e1A, e1B, 0x6240), e2 = __byte_perm(e2A, e2B, 0x6240);  // 4 bytes each
    a.e1s = __dp4a(e1, 0x01010101u, a.e1s);
    a.e1q = __dp4a(e1, e1, a.e1q);
    a.e2s = __dp4a(e2, 0x01010101u, a.e2s);
    a.e2q = __dp4a(e2, e2, a.e2q);
    // A5: Lap4 only where _detectLinearFeatures looks.  y0 and r0 are multiples of 4, so
    // y % 4 == i % 4; grid rows (i = 0, 4) and the rows below them (i = 1, 5) share a thread.
    if ((i & 3) <= 1) {
      const uint32_t crossA = up.cA + dn.cA + cur.hA;
      const uint32_t e3A = clip_diff(cur.cA * 5u, crossA);
      const uint32_t t0 = (e3A & 0xFFFFu) > IRP_SCRATCH_THRESHOLD, t1 = (e3A >> 16) > IRP_SCRATCH_THRESHOLD;
      if ((i & 3) == 0) {
        a.sv += FULL ? (t0 & t1) : (t0 & t1 & (uint32_t)(nvalid >= 2));
        prev_t0 = t0;
      } else {
        a.sh += prev_t0 & t0;
      }
    }
    // additive diagnostic: grey steps across 8-px boundaries
    if (col_boundary) {
      const int d = dp2a_lo_u16_s8(cur.p34, 0x0000FF01u, IRP_BLOCK_EDGE_THRESHOLD);  // c3 - r0 + T
      a.be0 += (uint32_t)d > 2u * IRP_BLOCK_EDGE_THRESHOLD;
    }
    if ((i & 7) == 7 && (FULL || y0 + r0 + i + 1 < H)) {
      // |a - b| > T per byte: saturating-free form on 16x2 halves
      const uint32_t ad = __vabsdiffu4(cur.word, dn.word);
#pragma unroll
      for (int j = 0; j < 4; j++)
        a.be1 += (uint32_t)(FULL || j < nvalid) & (uint32_t)(((ad >> (8 * j)) & 0xFFu) > IRP_BLOCK_EDGE_THRESHOLD);
    }
    up = cur;
    cur = dn;
  }
}

// ---------------------------------------------------------------------------
// stage 3: separable [12,20,12]/44 blur == floor((3(l+r)+5c+5)/11), H then V,
// u8 between passes; only pooled sum / sumsq of the result are kept.
// ---------------------------------------------------------------------------
// Division: floor(x / 11) for 0 <= x <= 2810 is byte 2 of x * 5958 (5958 * 11 = 2^16 + 2, and
// 2810 * 2 < 2^16; the product stays below 2^24), so a quotient never has to be shifted out: the
// next step picks it with PRMT.
struct HRow { uint32_t z[4]; };   // per column: (3(l+r) + 5c + 5) * 5958, the blurred byte is byte 2

// the blur's three constants: tap weights as dp4a bytes (edge, centre, edge, 0), rounding term, and the multiplier
// that leaves the quotient in byte 2.  IRP_BLUR_EXACT: (3, 5, 3) + 5, x 5958 (= / 11); IRP_BLUR_VECTOR (libvips'
// SIMD convi as recalled, 8-bit mantissas): (35, 58, 35) + 64, x 512 (= >> 7; 32704 * 512 < 2^24).
struct BlurK { uint32_t w, bias, mul; };
__host__ __device__ constexpr BlurK blur_exact() { return BlurK{0x00030503u, 5u, 5958u}; }
__host__ __device__ constexpr BlurK blur_vector() { return BlurK{0x00233A23u, 64u, 512u}; }

__device__ __forceinline__ HRow hpass_row(const uint8_t* plane_row, int strip, const BlurK bk = blur_exact()) {
  const uint32_t* rp = reinterpret_cast<const uint32_t*>(plane_row) + 3 + strip;
  const uint32_t l = rp[0], c = rp[1], r = rp[2];
  HRow o;
  const uint32_t w0 = __funnelshift_r(l, c, 24);  // (l3, c0, c1, c2)
  const uint32_t w3 = __funnelshift_r(c, r, 16);  // (c2, c3, r0, r1)
  o.z[0] = __dp4a(w0, bk.w, bk.bias) * bk.mul;
  o.z[1] = __dp4a(c, bk.w, bk.bias) * bk.mul;
  o.z[2] = __dp4a(c, bk.w << 8, bk.bias) * bk.mul;
  o.z[3] = __dp4a(w3, bk.w, bk.bias) * bk.mul;
  return o;
}

template <int C, bool FULL, int PITCH = kPlanePitch>
__device__ __forceinline__ void stage3(const Tiles<C>& T, Acc<C>& a, int tid, int x0, int y0, int W, int H, const BlurK bk = blur_exact()) {
  const int strip = tid & (kStrips - 1), rg = tid / kStrips;
  const int x = x0 + strip * 4;
  const int r0 = rg * 8;
  int nvalid = 4, nrows = 8;
  if (!FULL) {
    if (x >= W || y0 + r0 >= H) return;
    nvalid = min(4, W - x);
    nrows = min(8, H - (y0 + r0));
  }
  const uint32_t vmask = nvalid >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nvalid)) - 1u);
#pragma unroll 1   // one copy of the row loop: the hot code has to stay inside the instruction cache
  for (int ch = 0; ch < C; ch++) {
    const uint8_t* pl = T.plane[0] + ch * (kRows * PITCH);
    // per column a sliding window of horizontal results: bytes (up, cur, dn, 0)
    uint32_t win[4];
    {
      const HRow up = hpass_row(pl + r0 * PITCH, strip, bk), cur = hpass_row(pl + (r0 + 1) * PITCH, strip, bk);
#pragma unroll
      for (int j = 0; j < 4; j++) win[j] = __byte_perm(up.z[j], cur.z[j], 0x7620);  // (-, up, cur, 0)
    }
#pragma unroll 4
    for (int i = 0; i < 8; i++) {
      if (!FULL && i >= nrows) break;
      const HRow dn = hpass_row(pl + (r0 + i + 2) * PITCH, strip, bk);
      uint32_t zz[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        win[j] = __byte_perm(win[j], dn.z[j], 0x7621);                 // (up, cur, dn, 0)
        zz[j] = __dp4a(win[j], bk.w, bk.bias) * bk.mul;                // vertical result in byte 2
      }
      uint32_t b4 = __byte_perm(__byte_perm(zz[0], zz[1], 0x0062), __byte_perm(zz[2], zz[3], 0x0062), 0x5410);
      if (!FULL) b4 &= vmask;
      a.bs = __dp4a(b4, 0x01010101u, a.bs);
      a.bq = __dp4a(b4, b4, a.bq);
    }
  }
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kClassifyThreads, 3)
classify_kernel(const ImgDev* __restrict__ imgs, int n_imgs, int total_tiles, const ClassifyTables* __restrict__ tab,
                unsigned long long* __restrict__ gacc, uint32_t* __restrict__ ghist, uint32_t dyn_smem_bytes,
                int* __restrict__ error_flag, BlurK bk) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const SmemMap map = make_smem_map(sbase, C);
  if (map.end - sbase > dyn_smem_bytes) {  // the host sized the allocation from a probed base: must agree
    if (threadIdx.x == 0) atomicExch(error_flag, 1);
    return;
  }
  const int group = threadIdx.x / kGroupThreads, tid = threadIdx.x & (kGroupThreads - 1);
  Tiles<C> T;
  T.grey = smem_raw + ((group ? map.tiles[1] : map.tiles[0]) - sbase);
#pragma unroll
  for (int ch = 0; ch < C; ch++) T.plane[ch] = T.grey + (ch + 1) * kTileBytes;
  T.hist = reinterpret_cast<uint32_t*>(smem_raw + ((group ? map.hist[1] : map.hist[0]) - sbase));
  T.red = reinterpret_cast<uint32_t*>(smem_raw + ((group ? map.red[1] : map.red[0]) - sbase));
  T.a_lut_r = map.lut_r; T.a_lut_g = map.lut_g; T.a_lut_b = map.lut_b; T.a_inv = map.inv;
  T.a_hist = group ? map.hist[1] : map.hist[0];
  {
    uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw + (map.lut_r - sbase));
    uint32_t* inv = reinterpret_cast<uint32_t*>(smem_raw + (map.inv - sbase));
    for (int i = threadIdx.x; i < 3 * 256; i += kClassifyThreads) lut[i] = (&tab->lut[0][0])[i];
    for (int i = threadIdx.x; i < 4096; i += kClassifyThreads) inv[i] = tab->inv[i] - ((uint32_t)i << 20);
    for (int b = tid; b < 256; b += kGroupThreads) T.hist[b] = 0;
  }
  __syncthreads();

  // from here on the two groups never meet again: each walks its own tile sequence with its own barrier
  Acc<C> acc;
  acc.clear();
  int img = 0, cur_img = -1, since_flush = 0;
  for (int tile = blockIdx.x * kGroups + group; tile < total_tiles; tile += gridDim.x * kGroups) {
    while (img + 1 < n_imgs && tile >= imgs[img + 1].tile_base) img++;
    if (img != cur_img || since_flush >= kFlushTiles) {
      if (cur_img >= 0) {
        const int slot = imgs[cur_img].slot;
        flush_acc<C>(T, acc, gacc + (size_t)slot * ACC_COUNT, ghist + (size_t)slot * 256, tid, group);
      }
      cur_img = img;
      since_flush = 0;
    }
    since_flush++;
    const ImgDev im = imgs[img];
    const int t = tile - im.tile_base;
    const int ty = t / im.tiles_x, tx = t - ty * im.tiles_x;
    const int x0 = tx * kTileW, y0 = ty * kTileH;

    // ---- stage 1: load, regroup, moments, grey -> shared planar tiles ----
    // core rows (1..32): 256 sixteen-pixel items, exactly two per thread
    for (int item = tid; item < kTileH * kSegsPerRow; item += kGroupThreads) {
      const int row = 1 + item / kSegsPerRow, seg = item & (kSegsPerRow - 1);
      const int yy = y0 - 1 + row;
      const int gy = min(yy, im.h - 1);
      const bool counted = yy < im.h;
      const int xb = x0 + seg * kSegPx;
      if (xb >= im.w + 4) continue;  // right of the image and of every replicate column a strip reads
      bool fast = false;
      if constexpr (C == 3) {
        if (im.aligned16 && xb + kSegPx <= im.w) {
          const uint8_t* p = im.px + (size_t)gy * im.pitch + (size_t)xb * 3;
          if (counted)
            stage1_fast<true>(T, acc, p, row, 16 + seg * kSegPx);
          else
            stage1_fast<false>(T, acc, p, row, 16 + seg * kSegPx);
          fast = true;
        }
      }
      if (!fast) stage1_slow<C>(T, acc, im, gy, xb, row, 16 + seg * kSegPx, kSegPx, counted);
    }
    // fringe: the two halo rows and the two halo columns as single-pixel jobs spread over the whole
    // group, so no warp carries a third sixteen-pixel item to the barrier
    for (int f = tid; f < 2 * kTileW + 2 * kRows; f += kGroupThreads) {
      int row, xq, col;
      if (f < 2 * kTileW) {
        row = f < kTileW ? 0 : kRows - 1;
        const int px = f & (kTileW - 1);
        xq = x0 + px;
        col = 16 + px;
        if (xq >= im.w + 4) continue;
      } else {
        const int h = f - 2 * kTileW;
        row = h >> 1;
        xq = (h & 1) ? x0 + kTileW : x0 - 1;
        col = (h & 1) ? 16 + kTileW : 15;
      }
      const int gy = min(max(y0 - 1 + row, 0), im.h - 1);
      stage1_slow<C>(T, acc, im, gy, xq, row, col, 1, false);
    }
    group_barrier(group);
    // ---- L2 prefetch of this group's next tile (same image only) ----
    {
      constexpr int kLines = (kTileW * C + 127) / 128;
      const int tn = t + (int)gridDim.x * kGroups;
      if (tn < im.tiles_x * im.tiles_y && tid < kRows * kLines) {
        const int tyn = tn / im.tiles_x, txn = tn - tyn * im.tiles_x;
        const int row = tid / kLines, l = tid - row * kLines;
        const int gy = min(max(tyn * kTileH - 1 + row, 0), im.h - 1);
        const long long off = (long long)txn * kTileW * C + l * 128;
        if (off < (long long)im.w * C) {
          const uint8_t* pp = im.px + (size_t)gy * im.pitch + off;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pp));
        }
      }
    }
    // ---- stage 2 + 3 ----
    if (x0 + kTileW < im.w && y0 + kTileH < im.h) {
      stage2<C, true>(T, acc, tid, x0, y0, im.w, im.h);
      stage3<C, true>(T, acc, tid, x0, y0, im.w, im.h, bk);
    } else {
      stage2<C, false>(T, acc, tid, x0, y0, im.w, im.h);
      stage3<C, false>(T, acc, tid, x0, y0, im.w, im.h, bk);
    }
    group_barrier(group);
  }
  if (cur_img >= 0) {
    const int slot = imgs[cur_img].slot;
    flush_acc<C>(T, acc, gacc + (size_t)slot * ACC_COUNT, ghist + (size_t)slot * 256, tid, group);
  }
}

// the .shared address dynamic shared memory starts at, for a kernel without static shared memory
__global__ void probe_smem_base_kernel(uint32_t* out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  if (threadIdx.x == 0) *out = (uint32_t)__cvta_generic_to_shared(smem_raw);
}

}  // namespace irp
