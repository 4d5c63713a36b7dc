// irp_classify_bulk.cuh — the streaming variant of the fused degradation-statistics kernel for the
// common case (3-channel images whose base and pitch are multiples of 16 bytes: every host input
// once staged, and tight device inputs whose row length is a multiple of 16).
//
// Same arithmetic and the same stage 2 / stage 3 code as classify_kernel<3> (irp_classify.cuh; the
// reference formulas are cited there), different data movement:
//   * one CTA per SM, six independent 4-warp groups sharing one copy of the grey tables;
//   * each group owns a RAW tile buffer (34 rows x 416 bytes).  One lane of the group fills it with
//     ONE 2-D TMA tile load (cp.async.bulk.tensor -> UTMALDG, completion on an mbarrier) for tile
//     n+1 while the group runs the stencils of tile n, so stage 1 never waits on HBM: it reads the
//     interleaved bytes with LDS.128, and the halo rows / halo columns come out of the same buffer
//     instead of a byte-wise global path;
//   * the next tile's coordinates are worked out once, by the issuing warp, and handed over in
//     shared memory (no per-thread divisions, no per-thread descriptor reloads).
// The image is described to the TMA unit as a 2-D tensor of u16 elements (box 208 x 34); bytes
// outside the image arrive as zeros, and the replicate rows / columns of edge tiles are patched in
// the raw buffer before stage 1.
#pragma once
#include "irp_classify.cuh"

// Timing experiments only (tools/ablate.sh builds side libraries with -DIRP_ABLATE=mask; their RESULTS ARE WRONG):
// 1 no histogram atomics, 2 no stage 2, 4 no stage 3, 8 grey without the tables, 16 no channel moments.
#ifndef IRP_ABLATE
#define IRP_ABLATE 0
#endif

namespace irp {

constexpr int kBGroups = 6;                          // 4-warp groups per CTA
constexpr int kBThreads = kBGroups * kGroupThreads;  // 768
constexpr int kBPitch = 16 + kTileW;                 // plane row: [..15 = left halo][128 px]; right halo = byte 0 of the next row
constexpr int kBPlane = kRows * kBPitch;
constexpr int kBPlaneSet = 4 * kBPlane + 16;         // grey, R, G, B (+ the last row's right halo)
constexpr int kRawPitch = 16 + kTileW * 3 + 16;      // [..13-15 = left halo px][384 B][right halo px = 400-402 ..]
constexpr int kRawBytes = (kRows * kRawPitch + 127) & ~127;   // the TMA box (208 u16 x 34 rows), 128-byte aligned slots

struct __align__(64) TmaDesc { unsigned long long opaque[16]; };   // a CUtensorMap (cuda.h), 128 bytes

struct TileInfo {        // written by the issuing lane, read by the whole group
  int x0, y0, w, h;
  int slot;              // accumulator slot of the tile's image
  int flags;             // bit 0: tile valid (0 = end of this group's sequence); bit 1: flush the accumulators first
  int pad[2];
};

struct BulkMap {         // .shared window addresses
  uint32_t raw[kBGroups], planes[kBGroups], hist[kBGroups], red[kBGroups];
  uint32_t sync;         // per group: mbarrier (8 B) + 2 x TileInfo, 80 B stride
  uint32_t inv, lut_r, end;
};
constexpr int kSyncStride = 80;

__host__ __device__ inline BulkMap make_bulk_map(uint32_t base) {
  BulkMap m;
  uint32_t p = (base + 127u) & ~127u;
  m.raw[0] = p;
  p += kRawBytes;
  m.sync = (p + 15u) & ~15u;
  p = m.sync + kBGroups * kSyncStride;
  m.inv = (p + 16383u) & ~16383u;
  m.lut_r = m.inv + 16384u;
  p = m.lut_r + 3072u;
  for (int g = 0; g < kBGroups; g++) m.hist[g] = p + 1024u * g;
  p += 1024u * kBGroups;
  for (int g = 0; g < kBGroups; g++) m.red[g] = p + kRedBytes * g;
  p += kRedBytes * kBGroups;
  p = (p + 15u) & ~15u;
  for (int g = 0; g < kBGroups; g++) m.planes[g] = p + kBPlaneSet * g;
  p += kBPlaneSet * kBGroups;
  p = (p + 127u) & ~127u;
  for (int g = 1; g < kBGroups; g++) m.raw[g] = p + kRawBytes * (g - 1);
  p += kRawBytes * (kBGroups - 1);
  m.end = p;
  return m;
}

// ---- mbarrier / bulk-copy wrappers (shared::cta addresses) ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// one 2-D TMA tile load (box = the whole raw tile), completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_b32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// A multiplier the compiler must not see as a constant: a kernel parameter.  ptxas turns x * 4 + b
// with a literal 4 into LEA, which runs on the half-rate ALU pipe — the pipe that bounds this kernel;
// with an opaque operand it stays an IMAD on the FMA pipe.  (IMAD.HI as a right shift was measured
// too: slower than SHF + LOP3, so the inverse-table and histogram addresses keep those.)
struct MulConsts { uint32_t four; };

// Table addresses by IDP.4A: byte J of a packed word times 4 plus the table's base is ONE dot product with the
// selector (4 << 8J) — extraction, scaling and base add in a single instruction (it was PRMT / SHF + IMAD, and for the
// histogram SHF + LOP3), and it leaves the half-rate ALU pipe to the regroup and the stencils.
__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
template <int J>
__device__ __forceinline__ uint32_t grey_top_dp(const Tiles<3>& T, uint32_t R, uint32_t G, uint32_t B) {
  constexpr uint32_t sel = 4u << (8 * J);
  const uint32_t I = lds_u32(dp4a_uu(R, sel, T.a_lut_r)) + lds_u32(dp4a_uu(G, sel, T.a_lut_g)) + lds_u32(dp4a_uu(B, sel, T.a_lut_b));
  return lds_u32(((I >> 18) & 0x3FFCu) | T.a_inv) + I;
}
__device__ __forceinline__ void hist_add_dp(const Tiles<3>& T, uint32_t t) {
  red_inc_shared(dp4a_uu(t, 0x04000000u, T.a_hist));
}

// Horizontal [12,20,12]/44 pass (== floor((3(l+r) + 5c + 5) / 11), irp_classify.cuh stage 3) of the four pixels of planar
// word c; l ends with the pixel left of them (its byte 3), r starts with the pixel right of them (its byte 0).  Stage 1
// runs it while the pixels are in registers and stores the BLURRED rows: stage 3 then only has the vertical pass to do,
// reads one word instead of three per strip and row, and the rows a thread's window shares with its neighbour's are
// blurred once instead of twice.
__device__ __forceinline__ uint32_t hblur_word(uint32_t l, uint32_t c, uint32_t r) {
  constexpr BlurK bk = blur_exact();
  const uint32_t w0 = __funnelshift_r(l, c, 24);  // (l3, c0, c1, c2)
  const uint32_t w3 = __funnelshift_r(c, r, 16);  // (c2, c3, r0, r1)
  const uint32_t z0 = __dp4a(w0, bk.w, bk.bias) * bk.mul, z1 = __dp4a(c, bk.w, bk.bias) * bk.mul;
  const uint32_t z2 = __dp4a(c, bk.w << 8, bk.bias) * bk.mul, z3 = __dp4a(w3, bk.w, bk.bias) * bk.mul;
  return __byte_perm(__byte_perm(z0, z1, 0x0062), __byte_perm(z2, z3, 0x0062), 0x5410);   // the quotients are the products' byte 2
}

// 4 pixels (12 interleaved bytes in w0..w2) -> planar words + grey word.
// nvalid < 4 masks the moments / histogram (image edge); COUNTED = false for halo rows.
template <bool COUNTED, bool MASKED>
__device__ __forceinline__ void s1_quad(const Tiles<3>& T, const MulConsts& mc, Acc<3>& a, uint32_t w0, uint32_t w1, uint32_t w2, int nvalid,
                                        uint32_t& R, uint32_t& G, uint32_t& B, uint32_t& Y) {
  // w0 = R0 G0 B0 R1 | w1 = G1 B1 R2 G2 | w2 = B2 R3 G3 B3
  R = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);
  G = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);
  B = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);
  if (COUNTED && !(IRP_ABLATE & 16)) {
    uint32_t r = R, g = G, b = B;
    if (MASKED) {
      const uint32_t m = nvalid >= 4 ? 0xFFFFFFFFu : ((1u << (8 * max(nvalid, 0))) - 1u);
      r &= m; g &= m; b &= m;
    }
    a.s[0] = __dp4a(r, 0x01010101u, a.s[0]);
    a.s[1] = __dp4a(g, 0x01010101u, a.s[1]);
    a.s[2] = __dp4a(b, 0x01010101u, a.s[2]);
    a.q[0] = __dp4a(r, r, a.q[0]);
    a.q[1] = __dp4a(g, g, a.q[1]);
    a.q[2] = __dp4a(b, b, a.q[2]);
  }
  uint32_t t[4];
  if (IRP_ABLATE & 8) {
    Y = G;
    return;
  }
  t[0] = grey_top_dp<0>(T, R, G, B);
  t[1] = grey_top_dp<1>(T, R, G, B);
  t[2] = grey_top_dp<2>(T, R, G, B);
  t[3] = grey_top_dp<3>(T, R, G, B);
  if (COUNTED && !(IRP_ABLATE & 1)) {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (!MASKED || j < nvalid) hist_add_dp(T, t[j]);
  }
  Y = __byte_perm(__byte_perm(t[0], t[1], 0x0073), __byte_perm(t[2], t[3], 0x0073), 0x5410);
}

// one 16-pixel segment of a core row: 3 x LDS.128 (+ the two neighbour pixels) from the raw tile -> 4 x STS.128 to the
// planes: grey, and the horizontally blurred R, G, B
template <bool COUNTED, bool MASKED>
__device__ __forceinline__ void s1_segment(const Tiles<3>& T, const MulConsts& mc, Acc<3>& a, uint32_t raw_addr, uint32_t plane_addr, int nvalid) {
  const uint4 v0 = lds_v4(raw_addr), v1 = lds_v4(raw_addr + 16), v2 = lds_v4(raw_addr + 32);
  const uint32_t lw = lds_b32(raw_addr - 4), rw = lds_b32(raw_addr + 48);   // (x, R, G, B) of pixel -1 | (R, G, B, x) of pixel 16
  const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
  uint32_t R[4], G[4], B[4], Y[4];
#pragma unroll
  for (int k = 0; k < 4; k++)
    s1_quad<COUNTED, MASKED>(T, mc, a, w[3 * k], w[3 * k + 1], w[3 * k + 2], nvalid - 4 * k, R[k], G[k], B[k], Y[k]);
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(plane_addr), "r"(Y[0]), "r"(Y[1]), "r"(Y[2]), "r"(Y[3]) : "memory");
  if (IRP_ABLATE & 4) return;
  // left neighbours as words that END with the pixel (byte 3), right neighbours as words that START with it (byte 0)
  const uint32_t lR = lw << 16, lG = lw << 8, lB = lw, rR = rw, rG = rw >> 8, rB = rw >> 16;
  uint32_t H[4];
  H[0] = hblur_word(lR, R[0], R[1]); H[1] = hblur_word(R[0], R[1], R[2]); H[2] = hblur_word(R[1], R[2], R[3]); H[3] = hblur_word(R[2], R[3], rR);
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(plane_addr + kBPlane), "r"(H[0]), "r"(H[1]), "r"(H[2]), "r"(H[3]) : "memory");
  H[0] = hblur_word(lG, G[0], G[1]); H[1] = hblur_word(G[0], G[1], G[2]); H[2] = hblur_word(G[1], G[2], G[3]); H[3] = hblur_word(G[2], G[3], rG);
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(plane_addr + 2 * kBPlane), "r"(H[0]), "r"(H[1]), "r"(H[2]), "r"(H[3]) : "memory");
  H[0] = hblur_word(lB, B[0], B[1]); H[1] = hblur_word(B[0], B[1], B[2]); H[2] = hblur_word(B[1], B[2], B[3]); H[3] = hblur_word(B[2], B[3], rB);
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(plane_addr + 3 * kBPlane), "r"(H[0]), "r"(H[1]), "r"(H[2]), "r"(H[3]) : "memory");
}

// Stage 3 of the streaming kernel: the VERTICAL [12,20,12]/44 pass over rows stage 1 already blurred horizontally
// (same arithmetic and the same sums as stage3<> of irp_classify.cuh, which tests/test_gpu_paths.py holds it equal to).
template <bool FULL>
__device__ __forceinline__ void stage3v(const Tiles<3>& T, Acc<3>& a, int tid, int x0, int y0, int W, int H) {
  constexpr BlurK bk = blur_exact();
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
  for (int ch = 0; ch < 3; ch++) {
    const uint32_t* hp = reinterpret_cast<const uint32_t*>(T.plane[0] + ch * kBPlane + r0 * kBPitch) + 4 + strip;
    // per column a sliding window of horizontal results: bytes (up, cur, dn, 0)
    uint32_t win[4];
    {
      const uint32_t up = hp[0], cur = hp[kBPitch / 4];
      win[0] = __byte_perm(up, cur, 0x7400); win[1] = __byte_perm(up, cur, 0x7510);
      win[2] = __byte_perm(up, cur, 0x7620); win[3] = __byte_perm(up, cur, 0x7730);      // (-, up, cur, 0)
    }
#pragma unroll 8
    for (int i = 0; i < 8; i++) {
      if (!FULL && i >= nrows) break;
      const uint32_t dn = hp[(i + 2) * (kBPitch / 4)];
      uint32_t zz[4];
      win[0] = __byte_perm(win[0], dn, 0x7421); win[1] = __byte_perm(win[1], dn, 0x7521);
      win[2] = __byte_perm(win[2], dn, 0x7621); win[3] = __byte_perm(win[3], dn, 0x7721);  // (up, cur, dn, 0)
#pragma unroll
      for (int j = 0; j < 4; j++) zz[j] = __dp4a(win[j], bk.w, bk.bias) * bk.mul;         // vertical result in byte 2
      uint32_t b4 = __byte_perm(__byte_perm(zz[0], zz[1], 0x0062), __byte_perm(zz[2], zz[3], 0x0062), 0x5410);
      if (!FULL) b4 &= vmask;
      a.bs = __dp4a(b4, 0x01010101u, a.bs);
      a.bq = __dp4a(b4, b4, a.bq);
    }
  }
}

// the issuing warp: describe this group's next tile and start its row copies
struct Issuer {
  int next_tile, stride, img, last_img, since_flush;
};

// One lane describes the group's next tile and starts its load: a single TMA instruction brings the
// 34 x 416-byte box whose first column is 16 bytes left of the tile.  Whatever lies outside the
// image arrives as zeros and is replaced by replicate rows / columns before stage 1.
__device__ __forceinline__ void issue_tile(Issuer& is, const ImgDev* __restrict__ imgs, const TmaDesc* __restrict__ tmaps, int n_imgs,
                                           int total_tiles, uint32_t bar, uint32_t info_addr, uint32_t raw_addr) {
  const int tile = is.next_tile;
  is.next_tile += is.stride;
  if (tile >= total_tiles) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(info_addr + 20), "r"(0) : "memory");
    mbar_arrive(bar);
    return;
  }
  while (is.img + 1 < n_imgs && tile >= __ldg(&imgs[is.img + 1].tile_base)) is.img++;
  const ImgDev* im = imgs + is.img;
  const int w = __ldg(&im->w), h = __ldg(&im->h), tiles_x = __ldg(&im->tiles_x), slot = __ldg(&im->slot);
  const int t = tile - __ldg(&im->tile_base);
  const int ty = t / tiles_x, tx = t - ty * tiles_x;
  const int x0 = tx * kTileW, y0 = ty * kTileH;
  const bool fresh = is.img != is.last_img || is.since_flush >= kFlushTiles;
  const bool flush = is.last_img >= 0 && fresh;
  if (fresh) is.since_flush = 0;
  is.since_flush++;
  is.last_img = is.img;
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(info_addr), "r"(x0), "r"(y0), "r"(w), "r"(h) : "memory");
  asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(info_addr + 16), "r"(slot), "r"(1 | (flush ? 2 : 0)) : "memory");
  // generic-proxy accesses to the raw buffer (stage 1 reads, edge patches) were ordered before this
  // lane by the group barrier; order them before the async-proxy write of the next tile as well
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_arrive_expect_tx(bar, kRows * kRawPitch);
  tma_load_2d(raw_addr, tmaps + is.img, (x0 * 3 - 16) >> 1, y0 - 1, bar);   // coordinates in u16 elements, rows
}

__global__ void __launch_bounds__(kBThreads, 1)
classify_bulk_kernel(const ImgDev* __restrict__ imgs, const TmaDesc* __restrict__ tmaps, int n_imgs, int total_tiles, const ClassifyTables* __restrict__ tab,
                     unsigned long long* __restrict__ gacc, uint32_t* __restrict__ ghist, uint32_t dyn_smem_bytes,
                     int* __restrict__ error_flag, MulConsts mc) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const BulkMap map = make_bulk_map(sbase);
  if (map.end - sbase > dyn_smem_bytes) {
    if (threadIdx.x == 0) atomicExch(error_flag, 1);
    return;
  }
  const int group = threadIdx.x / kGroupThreads, tid = threadIdx.x & (kGroupThreads - 1);
  uint32_t a_raw = map.raw[0], a_planes = map.planes[0], a_hist = map.hist[0], a_red = map.red[0];
#pragma unroll
  for (int g = 1; g < kBGroups; g++)
    if (group == g) {
      a_raw = map.raw[g]; a_planes = map.planes[g]; a_hist = map.hist[g]; a_red = map.red[g];
    }
  const uint32_t a_bar = map.sync + group * kSyncStride, a_info = a_bar + 16;
  Tiles<3> T;
  T.grey = smem_raw + (a_planes - sbase);
#pragma unroll
  for (int ch = 0; ch < 3; ch++) T.plane[ch] = T.grey + (ch + 1) * kBPlane;
  T.hist = reinterpret_cast<uint32_t*>(smem_raw + (a_hist - sbase));
  T.red = reinterpret_cast<uint32_t*>(smem_raw + (a_red - sbase));
  T.a_lut_r = map.lut_r; T.a_lut_g = map.lut_r + 1024u; T.a_lut_b = map.lut_r + 2048u; T.a_inv = map.inv;
  T.a_hist = a_hist;
  {
    uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw + (map.lut_r - sbase));
    uint32_t* inv = reinterpret_cast<uint32_t*>(smem_raw + (map.inv - sbase));
    for (int i = threadIdx.x; i < 3 * 256; i += kBThreads) lut[i] = (&tab->lut[0][0])[i];
    for (int i = threadIdx.x; i < 4096; i += kBThreads) inv[i] = tab->inv[i] - ((uint32_t)i << 20);
    for (int b = tid; b < 256; b += kGroupThreads) T.hist[b] = 0;
    if (tid == 0) mbar_init(a_bar, 1);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  Issuer is;
  is.next_tile = blockIdx.x * kBGroups + group;
  is.stride = gridDim.x * kBGroups;
  is.img = 0; is.last_img = -1; is.since_flush = 0;
  if (tid == 0) issue_tile(is, imgs, tmaps, n_imgs, total_tiles, a_bar, a_info, a_raw);

  Acc<3> acc;
  acc.clear();
  int cur_slot = -1;
  for (uint32_t it = 0;; it++) {
    mbar_wait(a_bar, it & 1u);
    const uint32_t ia = a_info + (it & 1u) * 32u;
    int x0, y0, W, H, slot, flags;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(y0), "=r"(W), "=r"(H) : "r"(ia));
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(slot), "=r"(flags) : "r"(ia + 16));
    if (!(flags & 1)) break;
    if (flags & 2) flush_acc<3>(T, acc, gacc + (size_t)cur_slot * ACC_COUNT, ghist + (size_t)cur_slot * 256, tid, group);
    cur_slot = slot;
    const int vw = min(kTileW, W - x0);          // valid columns of this tile
    const bool full = x0 + kTileW < W && y0 + kTileH < H;

    // ---- replicate rows and columns of edge tiles, patched in the raw buffer (TMA filled them with zeros) ----
    if (y0 == 0 || y0 + kTileH >= H) {
      const int last = H - y0;                       // tile row holding image row H - 1
      for (int r = 0; r < kRows; r++) {
        const int sr = r == 0 ? (y0 == 0 ? 1 : 0) : min(r, last);
        if (sr != r && tid < kRawPitch / 4) {
          const uint32_t v = lds_b32(a_raw + sr * kRawPitch + tid * 4);
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(a_raw + r * kRawPitch + tid * 4), "r"(v) : "memory");
        }
      }
      group_barrier(group);
    }
    if (x0 == 0 || x0 + kTileW >= W) {
      if (tid < kRows) {
        const uint32_t rr = a_raw + tid * kRawPitch;
        if (x0 == 0)
          for (int k = 0; k < 3; k++) sts_u8(rr + 13 + k, lds_u8(rr + 16 + k));
        if (x0 + kTileW >= W)
          for (int k = 0; k < 3; k++) sts_u8(rr + 16 + 3 * vw + k, lds_u8(rr + 16 + 3 * (vw - 1) + k));
      }
      group_barrier(group);
    }

    // ---- stage 1: raw tile -> planar R/G/B + grey tiles, channel moments, histogram ----
    if (full) {
#pragma unroll 1
      for (int i = 0; i < 2; i++) {  // core rows 1..32: 256 segments, two per thread
        const int item = tid + i * kGroupThreads;
        const int row = 1 + (item >> 3), seg = item & 7;
        s1_segment<true, false>(T, mc, acc, a_raw + row * kRawPitch + 16 + seg * 48, a_planes + row * kBPitch + 16 + seg * 16, 16);
      }
    } else {
      for (int i = 0; i < 2; i++) {
        const int item = tid + i * kGroupThreads;
        const int row = 1 + (item >> 3), seg = item & 7;
        const int nvalid = vw - seg * 16;
        if (nvalid <= -1) continue;  // right of the replicate column: nothing reads it
        const uint32_t ra = a_raw + row * kRawPitch + 16 + seg * 48, pa = a_planes + row * kBPitch + 16 + seg * 16;
        if (y0 - 1 + row < H)
          s1_segment<true, true>(T, mc, acc, ra, pa, nvalid);
        else
          s1_segment<false, false>(T, mc, acc, ra, pa, 0);
      }
    }
    // halo rows (0 and 33) as 4-pixel pieces, halo columns (grey only: the vertical blur pass has no horizontal
    // neighbours) as single pixels: spread over the group
    if (tid >= 64) {
      const int q = tid - 64;
      const int row = (q >> 5) ? kRows - 1 : 0, c4 = q & 31;
      if (full || c4 * 4 <= vw) {
        const uint32_t ra = a_raw + row * kRawPitch + 16 + c4 * 12, pa = a_planes + row * kBPitch + 16 + c4 * 4;
        uint32_t R, G, B, Y;
        s1_quad<false, false>(T, mc, acc, lds_b32(ra), lds_b32(ra + 4), lds_b32(ra + 8), 0, R, G, B, Y);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(pa), "r"(Y) : "memory");
        if (!(IRP_ABLATE & 4)) {
          const uint32_t lw = lds_b32(ra - 4), rw = lds_b32(ra + 12);
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(pa + kBPlane), "r"(hblur_word(lw << 16, R, rw)) : "memory");
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(pa + 2 * kBPlane), "r"(hblur_word(lw << 8, G, rw >> 8)) : "memory");
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(pa + 3 * kBPlane), "r"(hblur_word(lw, B, rw >> 16)) : "memory");
        }
      }
    }
    if (tid < 2 * kRows) {
      const int row = tid >> 1, right = tid & 1;
      // left halo: raw bytes 13..15 -> plane byte 15; right halo (pixel vw of the tile): raw 16 + 3 vw -> next row's byte 0
      const uint32_t ra = a_raw + row * kRawPitch + (right ? 16 + 3 * vw : 13);
      const uint32_t pa = a_planes + (right ? row * kBPitch + 16 + vw : row * kBPitch + 15);
      const uint32_t r = lds_u8(ra), g = lds_u8(ra + 1), b = lds_u8(ra + 2);
      sts_u8(pa, grey_top_off(T, r << 2, g << 2, b << 2) >> 24);
    }
    group_barrier(group);

    // ---- the raw buffer is free: start the copies of this group's next tile ----
    if (tid == 0) issue_tile(is, imgs, tmaps, n_imgs, total_tiles, a_bar, a_info + ((it + 1) & 1u) * 32u, a_raw);

    // ---- stage 2 + 3 on the planes ----
    if (full) {
      if (!(IRP_ABLATE & 2)) stage2<3, true, kBPitch>(T, acc, tid, x0, y0, W, H);
      if (!(IRP_ABLATE & 4)) stage3v<true>(T, acc, tid, x0, y0, W, H);
    } else {
      if (!(IRP_ABLATE & 2)) stage2<3, false, kBPitch>(T, acc, tid, x0, y0, W, H);
      if (!(IRP_ABLATE & 4)) stage3v<false>(T, acc, tid, x0, y0, W, H);
    }
    group_barrier(group);
  }
  if (cur_slot >= 0) flush_acc<3>(T, acc, gacc + (size_t)cur_slot * ACC_COUNT, ghist + (size_t)cur_slot * 256, tid, group);
}

}  // namespace irp
