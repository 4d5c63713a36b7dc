// irp_resize_mma.cuh — lanczos3 resize (reference: server-node/src/middleware/imagePreprocess.js:46-53 = libvips
// reducev then reduceh, 12-bit fixed point, u8 between the passes) with both passes stated as banded integer
// matrix products on the 5th-generation tensor cores (tcgen05.mma kind::i8, int32 accumulators in tensor memory).
//
// Why: the ALU formulation (irp_resize_tma.cuh) is issue-bound — 13 taps x 2 passes of IDP.2A plus the loads
// that feed them cost ~45 thread-instructions per source pixel.  Here the multiply-accumulates leave the
// instruction stream: what the ALUs still do per byte is recombine / round / clip / pack (~3 instructions).
//
// Exactness.  A 12-bit signed coefficient does not fit the 8-bit operand, so every coefficient is split on the
// host as c = 128 * hi + lo with lo in [-64, 63] (both halves s8); the hi and lo products are accumulated in
// SEPARATE int32 columns of the same MMA (the coefficient operand is the N side: columns [0, 32) hold hi, columns
// [32, 64) lo, of the same 32 outputs) and recombined as 128 * acc_hi + acc_lo — integer arithmetic end to end,
// bit-identical to (sum c * p + 2048) >> 12.  The rounding constant 2048 is pre-loaded into the lo accumulator
// columns (tcgen05.st) and the MMAs accumulate on top of it.  Replicate edges are folded into the coefficient
// rows on the host (taps that fall outside the image are added to the edge tap), so tiles need no patching and
// TMA's zero fill outside the image is multiplied by zero.
//
// Tile = 128 output rows x 32 output columns.
//   * source footprint: rows x 256 bytes, loaded by TMA as two 128-byte column blocks with the 128-byte swizzle
//     (that IS the canonical MN-major operand layout, see tools/probes/umma_probe.cu), double buffered;
//   * reducev, per quarter of 32 output rows and per column block:  D[x byte (lane)][hi | lo of 32 rows] +=
//     SRC^T[x byte][source row] * CV[hi | lo row][source row]   (M 128, N 64, K 32 per step).
//     Epilogue: thread = one source byte column; 16 rows -> one 16-byte store into the PLANAR intermediate
//     (channel = byte index mod 3 is just part of the address): de-interleaving costs nothing;
//   * reduceh, per channel plane:  D[row (lane)][hi | lo of 32 output pixels] += MID_c[row][pixel] * CH[.][pixel];
//     epilogue: thread = one output row, packs R, G, B of 8 pixels into 24 interleaved bytes, three 8-byte stores.
//   * CV / CH (the banded coefficient matrices) are rebuilt only when the tile's pattern key changes: for the
//     usual ratios (4000 -> 2048 is 125 : 64) every tile of an image shares one pattern.
// One CTA per SM, 16 warps; one thread issues TMA and MMA; completion through mbarriers (tcgen05.commit).
#pragma once
#include "irp_classify_bulk.cuh"
#include "irp_resize.cuh"

namespace irp {

constexpr int kMmThreads = 512;
constexpr int kMmTR = 128, kMmTC = 32, kMmXB = 256;   // tile rows, tile columns, source bytes per tile row
constexpr int kMmMaxTaps = 32;                        // taps per output after edge folding (25 + slack)
constexpr int kMmKH = 96;                             // K of the horizontal pass, pixels (3 steps of 32)
constexpr int kMmMidPlane = (kMmTR / 16) * kMmKH * 16;   // bytes per channel plane of the intermediate
constexpr int kMmChBytes = 64 * kMmKH;
constexpr int kMmMaxKsv = 5;

struct MmJob {
  uint8_t* dst;
  unsigned long long dst_pitch;
  const int8_t* vtab;      // [dh][64]: hi[32] | lo[32], edge-folded taps of the output row
  const int32_t* vfirst;   // [dh] first source row the taps apply to (clamped into the image)
  const int32_t* vkey;     // [tiles_y] pattern id of the row block's CV
  const int8_t* htab;      // [dw][64]
  const int32_t* hfirst;   // [dw]
  const int32_t* hkey;     // [tiles_x]
  int sw, sh, dw, dh;
  int dst_x0, dst_y0;
  int tiles_x, tiles_y, tile_base;
  int vnt, hnt;            // taps per output row / column
  int ksv;                 // 32-row K steps per quarter
  int pad;
};

struct MmLayout {          // byte offsets in dynamic shared memory (after 1024-byte alignment)
  int R;                   // rows of a source buffer (multiple of 16)
  int ksv_max;
  int off_cv, off_mid, off_ch, off_info, off_bar, off_src, total;
};

struct MmInfo {            // one tile, written by the issuing thread
  int job, ox0, oy0, sy0;
  int bx0, ws0, ws1, ws2;
  int ws3, vkey, hkey, valid;
};

__device__ __forceinline__ unsigned long long mm_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (unsigned long long)((saddr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo >> 4) & 0x3FFFu) << 16) |
         ((unsigned long long)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((unsigned long long)layout << 61);
}
__device__ __forceinline__ void mm_mma_i8(uint32_t d, unsigned long long a, unsigned long long b, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void mm_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mm_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
}
__device__ __forceinline__ void mm_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void mm_st16(uint32_t taddr, uint32_t v) {   // 16 columns of this warp's 32 lanes <- v
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void mm_st8(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void mm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mm_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// (128 * hi + lo) >> 12, four of them saturated to bytes; the accumulators already hold the rounding constant
__device__ __forceinline__ uint32_t mm_pack4(const uint32_t* hi, const uint32_t* lo) {
  int v[4];
#pragma unroll
  for (int j = 0; j < 4; j++) v[j] = ((int)hi[j] * 128 + (int)lo[j]) >> IRP_INTERP_SHIFT;
  return pack_sat_u8(v[1], v[0], pack_sat_u8(v[3], v[2], 0u));
}

constexpr uint32_t kMmIdesc = (2u << 4) /* s32 */ | (0u << 7) /* A u8 */ | (1u << 10) /* B s8 */ | (1u << 15) /* A MN-major */ |
                              ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kMmColV = 0, kMmColH = 256;   // accumulator columns: reducev 2 x 128, reduceh 3 x 64

// the issuing thread: describe tile `tile` and start the loads of its source footprint into buffer `buf`
__device__ __forceinline__ void mm_issue_tile(int tile, int tile_end, int& job, const MmJob* __restrict__ jobs, const TmaDesc* __restrict__ tmaps, int n_jobs,
                                              const MmLayout& L, uint32_t sbase, int buf) {
  const uint32_t info = sbase + L.off_info + buf * (uint32_t)sizeof(MmInfo), bar = sbase + L.off_bar + 8u * buf;
  if (tile >= tile_end) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(info + 44), "r"(0) : "memory");
    mbar_arrive(bar);
    return;
  }
  while (job + 1 < n_jobs && tile >= __ldg(&jobs[job + 1].tile_base)) job++;
  const MmJob* J = jobs + job;
  const int tiles_y = __ldg(&J->tiles_y), dh = __ldg(&J->dh);
  const int t = tile - __ldg(&J->tile_base);
  const int strip = t / tiles_y, rb = t - strip * tiles_y;
  const int ox0 = strip * kMmTC, oy0 = rb * kMmTR;
  const int32_t* vfirst = reinterpret_cast<const int32_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&J->vfirst)));
  const int32_t* hfirst = reinterpret_cast<const int32_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&J->hfirst)));
  const int32_t* vkey = reinterpret_cast<const int32_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&J->vkey)));
  const int32_t* hkey = reinterpret_cast<const int32_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&J->hkey)));
  const int sy0 = __ldg(vfirst + oy0) & ~7;
  int ws[4];
#pragma unroll
  for (int q = 0; q < 4; q++) ws[q] = (__ldg(vfirst + min(oy0 + 32 * q, dh - 1)) & ~7) - sy0;
  const int bx0 = (3 * __ldg(hfirst + ox0)) & ~15;
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(info), "r"(job), "r"(ox0), "r"(oy0), "r"(sy0) : "memory");
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(info + 16), "r"(bx0), "r"(ws[0]), "r"(ws[1]), "r"(ws[2]) : "memory");
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(info + 32), "r"(ws[3]), "r"(__ldg(vkey + rb)), "r"(__ldg(hkey + strip)), "r"(1) : "memory");
  const uint32_t src = sbase + L.off_src + (uint32_t)buf * (2u * L.R * 128u);
  mbar_arrive_expect_tx(bar, 2u * L.R * 128u);
#pragma unroll
  for (int b = 0; b < 2; b++)
#pragma unroll
    for (int h = 0; h < 2; h++)
      tma_load_2d(src + b * (L.R * 128) + h * (L.R / 2) * 128, tmaps + job, bx0 + 128 * b, sy0 + h * (L.R / 2), bar);
}

__global__ void __launch_bounds__(kMmThreads, 1)
resize_mma_kernel(const MmJob* __restrict__ jobs, const TmaDesc* __restrict__ tmaps, int n_jobs, int total_tiles, int tiles_per_cta, MmLayout L,
                  long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t mm_smem[];
  __shared__ uint32_t s_tmem;
  const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(mm_smem) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_cv = sbase + L.off_cv, a_mid = sbase + L.off_mid, a_ch = sbase + L.off_ch, a_info = sbase + L.off_info,
                 a_bar = sbase + L.off_bar, a_src = sbase + L.off_src;
  // mbarriers: [0,1] source buffers, [2,3] reducev accumulators, [4] reduceh accumulator
  if (tid == 0)
    for (int k = 0; k < 5; k++) mbar_init(a_bar + 8 * k, 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mm_fence_before();
  __syncthreads();
  mm_fence_after();
  const uint32_t tm = s_tmem;
  const int lq = warp & 3;
  const uint32_t tlane = tm + ((uint32_t)(32 * lq) << 16);
  // V epilogue item of this warp: column block b, rows [16 rh, 16 rh + 16) of the quarter
  const int vb = (warp >> 2) & 1, vrh = warp >> 3;
  // H epilogue item: pixels [8 pg, 8 pg + 8) of the tile's 32 output columns
  const int pg = warp >> 2;
  // arm every accumulator: hi columns 0, lo columns 2048 (the rounding constant)
#pragma unroll
  for (int dv = 0; dv < 2; dv++) {
    mm_st16(tlane + kMmColV + dv * 128 + vb * 64 + 16 * vrh, 0u);
    mm_st16(tlane + kMmColV + dv * 128 + vb * 64 + 32 + 16 * vrh, 1u << (IRP_INTERP_SHIFT - 1));
  }
#pragma unroll
  for (int c = 0; c < 3; c++) {
    mm_st8(tlane + kMmColH + 64 * c + 8 * pg, 0u);
    mm_st8(tlane + kMmColH + 64 * c + 32 + 8 * pg, 1u << (IRP_INTERP_SHIFT - 1));
  }
  mm_wait_st();

  const int tile_begin = blockIdx.x * tiles_per_cta, tile_end = min(total_tiles, tile_begin + tiles_per_cta);
  if (tile_begin >= tile_end) {
    mm_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    return;
  }
  int issue_job = 0;
  if (tid == 0) {
    mm_issue_tile(tile_begin, tile_end, issue_job, jobs, tmaps, n_jobs, L, sbase, 0);
    mm_issue_tile(tile_begin + 1, tile_end, issue_job, jobs, tmaps, n_jobs, L, sbase, 1);
  }
  int cur_vkey = -1, cur_hkey = -1, ch_sel = 0;
  const int cvq = L.ksv_max * 2048;   // bytes of one quarter's CV

  // all threads: (re)build the coefficient matrices of the tile in `inf` when its pattern differs from what is loaded
  auto build = [&](const MmInfo& inf, const MmJob* J) {
    const bool need_v = inf.vkey != cur_vkey, need_h = inf.hkey != cur_hkey;
    if (!need_v && !need_h) return;
    if (need_h) ch_sel ^= 1;
    const uint32_t chb = a_ch + ch_sel * kMmChBytes;
    if (need_v)
      for (int i = tid; i < 4 * cvq / 16; i += kMmThreads) mm_sts128(a_cv + 16 * i, 0, 0, 0, 0);
    if (need_h)
      for (int i = tid; i < kMmChBytes / 16; i += kMmThreads) mm_sts128(chb + 16 * i, 0, 0, 0, 0);
    __syncthreads();
    if (need_v) {
      const int r = tid >> 2, hl = (tid >> 1) & 1, th = tid & 1, oy = inf.oy0 + r;
      if (oy < J->dh) {
        const int q = r >> 5, n = hl * 32 + (r & 31);
        const int wsq = q == 0 ? inf.ws0 : (q == 1 ? inf.ws1 : (q == 2 ? inf.ws2 : inf.ws3));
        const int koff = __ldg(J->vfirst + oy) - (inf.sy0 + wsq);
        const int8_t* src = J->vtab + (size_t)oy * 64 + hl * 32;
        const uint32_t rowa = a_cv + q * cvq + (n >> 3) * 128 + (n & 7) * 16;
        const int jend = min(J->vnt, th * 16 + 16);
        for (int j = th * 16; j < jend; j++) {
          const int k = koff + j;
          sts_u8(rowa + (k >> 4) * 1024 + (k & 15), (uint32_t)(uint8_t)__ldg(src + j));
        }
      }
      cur_vkey = inf.vkey;
    }
    if (need_h) {
      const int r = tid >> 4, hl = (tid >> 3) & 1, part = tid & 7, ox = inf.ox0 + r;
      if (ox < J->dw) {
        const int n = hl * 32 + r;
        const int koff = __ldg(J->hfirst + ox) - inf.bx0 / 3;
        const int8_t* src = J->htab + (size_t)ox * 64 + hl * 32;
        const uint32_t rowa = chb + (n >> 3) * 128 + (n & 7) * 16;
        const int jend = min(J->hnt, part * 4 + 4);
        for (int j = part * 4; j < jend; j++) {
          const int k = koff + j;
          sts_u8(rowa + (k >> 4) * 1024 + (k & 15), (uint32_t)(uint8_t)__ldg(src + j));
        }
      }
      cur_hkey = inf.hkey;
    }
  };
  auto load_info = [&](int buf) {
    MmInfo inf;
    const uint32_t ia = a_info + buf * (uint32_t)sizeof(MmInfo);
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inf.job), "=r"(inf.ox0), "=r"(inf.oy0), "=r"(inf.sy0) : "r"(ia));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inf.bx0), "=r"(inf.ws0), "=r"(inf.ws1), "=r"(inf.ws2) : "r"(ia + 16));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inf.ws3), "=r"(inf.vkey), "=r"(inf.hkey), "=r"(inf.valid) : "r"(ia + 32));
    return inf;
  };
  // the issuing thread: the reducev MMAs of quarter q of the tile in source buffer `buf`
  auto issue_v = [&](const MmInfo& inf, int buf, int q, int ksv) {
    const int wsq = q == 0 ? inf.ws0 : (q == 1 ? inf.ws1 : (q == 2 ? inf.ws2 : inf.ws3));
    const uint32_t src = a_src + (uint32_t)buf * (2u * L.R * 128u);
#pragma unroll
    for (int b = 0; b < 2; b++)
      for (int ks = 0; ks < ksv; ks++) {
        const unsigned long long ad = mm_desc(src + b * (L.R * 128) + (wsq + 32 * ks) * 128, (uint32_t)L.R * 128u, 1024u, 2u);
        const unsigned long long bd = mm_desc(a_cv + q * cvq + ks * 2048, 1024u, 128u, 0u);
        mm_mma_i8(tm + kMmColV + (q & 1) * 128 + b * 64, ad, bd, kMmIdesc);
      }
    mm_commit(a_bar + 8 * (2 + (q & 1)));
  };

  // ---- prologue: first tile's matrices and its first two quarters ----
  MmInfo cur;
  {
    mbar_wait(a_bar, 0);
    cur = load_info(0);
    build(cur, jobs + cur.job);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      mm_fence_after();
      const int ksv = jobs[cur.job].ksv;
      issue_v(cur, 0, 0, ksv);
      issue_v(cur, 0, 1, ksv);
    }
  }
  int dbg_n = 0;
  auto stamp = [&](int it_) {
    if (dbg && tid == 0 && blockIdx.x == 0 && it_ == 3 && dbg_n < 32) dbg[dbg_n++] = clock64();
  };
  for (int it = 0;; it++) {
    const int buf = it & 1;
    const MmJob* J = jobs + cur.job;
    stamp(it);
    const int ksv = __ldg(&J->ksv);
    // planar address of this thread's source byte column (reducev epilogue)
    uint32_t mid_col;
    {
      const int xb = 128 * vb + 32 * lq + lane, B = cur.bx0 + xb, p = B / 3, c = B - 3 * p, kpx = p - cur.bx0 / 3;
      mid_col = a_mid + c * kMmMidPlane + (kpx >> 3) * 128 + (kpx & 7) * 16;
    }
    // ---- reducev: four quarters through two accumulator buffers ----
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
      mbar_wait(a_bar + 8 * (2 + (q & 1)), (uint32_t)(q >> 1) & 1u);
      mm_fence_after();
      stamp(it);
      uint32_t hi[16], lo[16];
      const uint32_t ta = tlane + kMmColV + (q & 1) * 128 + vb * 64 + 16 * vrh;
      mm_ld16(ta, hi);
      mm_ld16(ta + 32, lo);
      mm_wait_ld();
      mm_st16(ta, 0u);   // re-arm the accumulator for the quarter after next
      mm_st16(ta + 32, 1u << (IRP_INTERP_SHIFT - 1));
      mm_sts128(mid_col + (2 * q + vrh) * (kMmKH * 16), mm_pack4(hi, lo), mm_pack4(hi + 4, lo + 4), mm_pack4(hi + 8, lo + 8), mm_pack4(hi + 12, lo + 12));
      mm_wait_st();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the intermediate is read by the tensor core next
      mm_fence_before();
      __syncthreads();
      stamp(it);
      if (tid == 0 && q + 2 < 4) {
        mm_fence_after();
        issue_v(cur, buf, q + 2, ksv);
      }
    }
    // ---- reduceh of this tile ----
    if (tid == 0) {
      mm_fence_after();
      const uint32_t chb = a_ch + ch_sel * kMmChBytes;
#pragma unroll
      for (int c = 0; c < 3; c++)
#pragma unroll
        for (int ks = 0; ks < kMmKH / 32; ks++) {
          const unsigned long long ad = mm_desc(a_mid + c * kMmMidPlane + ks * 512, 128u, (uint32_t)kMmKH * 16u, 0u);
          const unsigned long long bd = mm_desc(chb + ks * 2048, 1024u, 128u, 0u);
          mm_mma_i8(tm + kMmColH + 64 * c, ad, bd, kMmIdesc);
        }
      mm_commit(a_bar + 8 * 4);
    }
    stamp(it);
    // ---- while it runs: the next tile's matrices, its first two quarters, and the loads of the tile after it ----
    mbar_wait(a_bar + 8 * (buf ^ 1), (uint32_t)((it + 1) >> 1) & 1u);
    const MmInfo nxt = load_info(buf ^ 1);
    const int my_ch = ch_sel;   // the matrix the reduceh in flight reads
    if (nxt.valid) {
      build(nxt, jobs + nxt.job);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        mm_fence_after();
        const int ksv2 = jobs[nxt.job].ksv;
        issue_v(nxt, buf ^ 1, 0, ksv2);
        issue_v(nxt, buf ^ 1, 1, ksv2);
        mm_issue_tile(tile_begin + it + 2, tile_end, issue_job, jobs, tmaps, n_jobs, L, sbase, buf);   // this tile's buffer is free
      }
    }
    (void)my_ch;
    stamp(it);
    // ---- reduceh epilogue: thread = output row, 8 pixels x RGB -> 24 interleaved bytes ----
    mbar_wait(a_bar + 8 * 4, (uint32_t)it & 1u);
    mm_fence_after();
    stamp(it);
    {
      uint32_t hi[3][8], lo[3][8];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        mm_ld8(tlane + kMmColH + 64 * c + 8 * pg, hi[c]);
        mm_ld8(tlane + kMmColH + 64 * c + 32 + 8 * pg, lo[c]);
      }
      mm_wait_ld();
#pragma unroll
      for (int c = 0; c < 3; c++) {
        mm_st8(tlane + kMmColH + 64 * c + 8 * pg, 0u);
        mm_st8(tlane + kMmColH + 64 * c + 32 + 8 * pg, 1u << (IRP_INTERP_SHIFT - 1));
      }
      int v[24];   // interleaved: v[3 j + c]
#pragma unroll
      for (int j = 0; j < 8; j++)
#pragma unroll
        for (int c = 0; c < 3; c++) v[3 * j + c] = ((int)hi[c][j] * 128 + (int)lo[c][j]) >> IRP_INTERP_SHIFT;
      uint32_t w[6];
#pragma unroll
      for (int k = 0; k < 6; k++) w[k] = pack_sat_u8(v[4 * k + 1], v[4 * k], pack_sat_u8(v[4 * k + 3], v[4 * k + 2], 0u));
      const int row = 32 * lq + lane, oy = cur.oy0 + row, px0 = cur.ox0 + 8 * pg;
      if (oy < J->dh && px0 < J->dw) {
        uint8_t* d = J->dst + (size_t)(J->dst_y0 + oy) * J->dst_pitch + (size_t)(J->dst_x0 + px0) * 3;
        if (px0 + 8 <= J->dw) {
#pragma unroll
          for (int k = 0; k < 3; k++) *reinterpret_cast<uint2*>(d + 8 * k) = make_uint2(w[2 * k], w[2 * k + 1]);
        } else {
          const int nb = 3 * (J->dw - px0);
          for (int k = 0; k < nb; k++) d[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
        }
      }
      mm_wait_st();
    }
    mm_fence_before();
    __syncthreads();
    stamp(it);
    if (!nxt.valid) break;
    cur = nxt;
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

}  // namespace irp
