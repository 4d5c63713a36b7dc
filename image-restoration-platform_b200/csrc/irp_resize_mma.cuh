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
// bit-identical to (sum c * p + 2048) >> 12.  The first K step of every accumulation overwrites its accumulator; the
// rounding constant is added in the reducev epilogue and, for reduceh, comes out of the product itself (a constant-1
// pixel column of the intermediate times a hi coefficient of 16).  Replicate edges are folded into the coefficient
// rows on the host (taps that fall outside the image are added to the edge tap), so tiles need no patching and
// TMA's zero fill outside the image is multiplied by zero.
//
// Tile = 128 output rows x 32 output columns; a CTA takes whole 32-column STRIPS from a queue and walks them down.
//   * source footprint: rows x 256 bytes, loaded by TMA as two 128-byte column blocks with the 128-byte swizzle
//     (that IS the canonical MN-major operand layout, see tools/probes/umma_probe.cu), double buffered; the box
//     starts at the first source row the tile needs, so the quarter windows sit at the same offsets in every
//     tile of a periodic geometry (4000 -> 2048 is 125 : 64);
//   * reducev, per quarter of 32 output rows and per column block:  D[x byte (lane)][hi | lo of 32 rows] +=
//     SRC^T[x byte][source row] * CV[hi | lo row][source row]   (M 128, N 64, K 32 per step).
//     Epilogue: thread = one source byte column; 16 rows -> one 16-byte store into the PLANAR intermediate
//     (channel = byte index mod 3 is just part of the address): de-interleaving costs nothing;
//   * reduceh, per channel plane:  D[row (lane)][hi | lo of 32 output pixels] += MID_c[row][pixel] * CH[.][pixel];
//     epilogue: thread = one output row, packs R, G, B of 8 pixels into 24 interleaved bytes, three 8-byte stores.
//   * CV / CH (the banded coefficient matrices, already in the operand layout) are built ON THE HOST per
//     (row block, quarter) / per strip, de-duplicated, and fetched by bulk copies into their shared-memory slots
//     only when the slot holds a different matrix.  (The first version rebuilt them in the kernel from the tap
//     tables: 8000 cycles of dependent L2 loads per tile.)
// Roles (one CTA per SM, 20 warps): warps 0-15 epilogues (tensor memory -> registers -> shared / global memory),
// warp 16 the producer (strip queue, tile records, TMA and bulk copies), warps 17-19 issue the MMAs (one lane each):
// the two column blocks of reducev and the three planes of reduceh come from different warps.
// Everything is handed over through mbarriers (tcgen05.commit on the tensor side); the tensor pipe sees, per tile,
//   H(t-1)  V1(t)  V2(t)  V3(t)  V0(t+1)
// and the epilogue warps run  EV0(t)  EH(t-1)  EV1(t)  EV2(t)  EV3(t), so that the reduceh of a tile overlaps the
// first reducev epilogue of the next one and the single intermediate buffer never idles the tensor pipe.
#pragma once
#include "irp_classify_bulk.cuh"
#include "irp_resize.cuh"

namespace irp {

constexpr int kMmEpiWarps = 16;
constexpr int kMmThreads = 32 * (kMmEpiWarps + 4);
constexpr int kMmVLanes = 2, kMmHLanes = 3;   // arrivals on the accumulator-complete barriers: one per column block of reducev / per plane of reduceh
constexpr int kMmTR = 128, kMmTC = 32, kMmXB = 256;   // tile rows, tile columns, source bytes per tile row
constexpr int kMmMaxTaps = 32;                        // taps per output after edge folding (25 + slack)
constexpr int kMmKH = 96;                             // K of the horizontal pass, pixels (3 steps of 32)
constexpr int kMmMidPlane = (kMmTR / 16) * kMmKH * 16;   // bytes per channel plane of the intermediate
constexpr int kMmChBytes = 64 * kMmKH;
constexpr int kMmMaxKsv = 5;
constexpr int kMmOutPitch = 104;                      // bytes per row of the output staging tile (96 + pad: 8-byte stores without bank conflicts)

struct MmJob {
  uint8_t* dst;
  unsigned long long dst_pitch;
  const int32_t* vfirst;   // [dh] first source row the (edge-folded) taps of an output row apply to
  const int32_t* vmat;     // [tiles_y * 4] matrix of the (row block, quarter)
  const uint8_t* vmats;    // matrices in the operand layout, ksv * 2048 bytes each
  const int32_t* hfirst;   // [dw]
  const int32_t* hmat;     // [tiles_x]
  const uint8_t* hmats;    // kMmChBytes each
  int sw, sh, dw, dh;
  int dst_x0, dst_y0;
  int tiles_x, tiles_y, strip_base;
  int ksv;                 // 32-row K steps per quarter
  int pad[2];
};

struct MmLayout {          // byte offsets in dynamic shared memory (after 1024-byte alignment)
  int R;                   // rows of a source buffer (multiple of 16)
  int ksv_max;
  int nbuf;                // source buffers: 2, or 1 for footprints that do not fit twice (shrinks above ~2)
  int off_cv, off_mid, off_ch, off_out, off_info, off_bar, off_src, total;
};

struct MmInfo {            // one tile, written by the producer
  int rows, cols, bx0, valid;   // output rows / columns of the tile inside the image; first source byte column
  int ws0, ws1, ws2, ws3;       // first source row of each quarter's window, relative to the box (multiples of 8)
  int flags, ksv;               // flags: bits 0-3 quarter matrix reloaded, bit 4 CH reloaded, bit 5 CH slot
  unsigned dst_lo, dst_hi;      // where the tile's first output byte goes
  unsigned long long pitch;
  int pad[2];
};

// mbarrier indices
enum {
  kBarFull = 0,      // [2] source buffer + tile record (producer, transaction bytes)
  kBarSrcFree = 2,   // [2] reducev of the tile is through with the source buffer (commits of the issuing lanes)
  kBarCvFull = 4,    // [4] quarter matrix landed
  kBarCvFree = 8,    // [4] quarter's MMAs done
  kBarChFull = 12,   // [2]
  kBarChFree = 14,   // [2]
  kBarVFull = 16,    // [2] reducev accumulators complete
  kBarVFree = 18,    // [2] ... read out and re-armed (16 epilogue warps)
  kBarHFull = 20,
  kBarHFree = 21,
  kBarMidFull = 22,  // intermediate tile written (16 epilogue warps)
  kBarCount = 23
};

__device__ __forceinline__ unsigned long long mm_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (unsigned long long)((saddr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo >> 4) & 0x3FFFu) << 16) |
         ((unsigned long long)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((unsigned long long)layout << 61);
}
__device__ __forceinline__ void mm_mma_i8(uint32_t d, unsigned long long a, unsigned long long b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void mm_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// (128 * hi + lo + 2048) >> 12, four of them saturated to bytes
__device__ __forceinline__ uint32_t mm_pack4(const uint32_t* hi, const uint32_t* lo) {
  int v[4];
#pragma unroll
  for (int j = 0; j < 4; j++) v[j] = ((int)hi[j] * 128 + ((int)lo[j] + (1 << (IRP_INTERP_SHIFT - 1)))) >> IRP_INTERP_SHIFT;
  return pack_sat_u8(v[1], v[0], pack_sat_u8(v[3], v[2], 0u));
}

constexpr uint32_t kMmIdesc = (2u << 4) /* s32 */ | (0u << 7) /* A u8 */ | (1u << 10) /* B s8 */ | (1u << 15) /* A MN-major */ |
                              ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kMmColV = 0, kMmColH = 256;   // accumulator columns: reducev 2 x 128, reduceh 3 x 64

__device__ __forceinline__ MmInfo mm_load_info(uint32_t ia) {
  MmInfo inf;
  unsigned plo, phi;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inf.rows), "=r"(inf.cols), "=r"(inf.bx0), "=r"(inf.valid) : "r"(ia));
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inf.ws0), "=r"(inf.ws1), "=r"(inf.ws2), "=r"(inf.ws3) : "r"(ia + 16));
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inf.flags), "=r"(inf.ksv), "=r"(inf.dst_lo), "=r"(inf.dst_hi) : "r"(ia + 32));
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(plo), "=r"(phi) : "r"(ia + 48));
  inf.pitch = ((unsigned long long)phi << 32) | plo;
  return inf;
}
template <typename T>
__device__ __forceinline__ const T* mm_ldptr(const T* const* p) {
  return reinterpret_cast<const T*>(__ldg(reinterpret_cast<const unsigned long long*>(p)));
}
__device__ __forceinline__ uint32_t mm_clock() {
  uint32_t c;
  asm volatile("mov.u32 %0, %%clock;" : "=r"(c));
  return c;
}

__global__ void __launch_bounds__(kMmThreads, 1)
resize_mma_kernel(const MmJob* __restrict__ jobs, const TmaDesc* __restrict__ tmaps, int n_jobs, int total_strips, int* __restrict__ strip_counter,
                  MmLayout L, long long* __restrict__ dbg) {
  extern __shared__ __align__(1024) uint8_t mm_smem[];
  __shared__ uint32_t s_tmem;
  const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(mm_smem) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_cv = sbase + L.off_cv, a_mid = sbase + L.off_mid, a_ch = sbase + L.off_ch, a_out = sbase + L.off_out, a_info = sbase + L.off_info,
                 a_bar = sbase + L.off_bar, a_src = sbase + L.off_src;
  auto bar = [&](int k) { return a_bar + 8u * (uint32_t)k; };
  // kernel experiments (IRP_MMA_DEBUG): cycles block 0 spends in each kind of wait and in the epilogue steps
  uint32_t wacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const bool prof = dbg != nullptr && blockIdx.x == 0;
  uint32_t tk = 0;
  auto tick = [&](int slot) {
    if (prof) {
      const uint32_t t = mm_clock();
      wacc[slot] += t - tk;
      tk = t;
    }
  };
  auto twait = [&](int slot, uint32_t b, uint32_t parity) {
    if (prof) {
      const uint32_t t0 = mm_clock();
      mbar_wait(b, parity);
      wacc[slot] += mm_clock() - t0;
    } else {
      mbar_wait(b, parity);
    }
  };
  const uint32_t t_start = mm_clock();
  if (tid == 0) {
    for (int k = 0; k < 2; k++) mbar_init(bar(kBarFull + k), 1), mbar_init(bar(kBarSrcFree + k), kMmVLanes);
    for (int k = 0; k < 4; k++) mbar_init(bar(kBarCvFull + k), 1), mbar_init(bar(kBarCvFree + k), kMmVLanes);
    for (int k = 0; k < 2; k++) mbar_init(bar(kBarChFull + k), 1), mbar_init(bar(kBarChFree + k), 1);
    for (int k = 0; k < 2; k++) mbar_init(bar(kBarVFull + k), kMmVLanes), mbar_init(bar(kBarVFree + k), kMmEpiWarps);
    mbar_init(bar(kBarHFull), kMmHLanes);
    mbar_init(bar(kBarHFree), kMmEpiWarps);
    mbar_init(bar(kBarMidFull), kMmEpiWarps);
  }
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
  // V epilogue item of an epilogue warp: column block vb, rows [16 vrh, 16 vrh + 16) of the quarter
  const int vb = (warp >> 2) & 1, vrh = warp >> 3;
  // H epilogue item: pixels [8 pg, 8 pg + 8) of the tile's 32 output columns
  const int pg = warp >> 2;
  if (warp < kMmEpiWarps) {
    // reduceh's rounding constant comes out of the product itself: pixel 95 of every intermediate row is a constant 1
    // (the reducev epilogue never writes past pixel 85) and the hi half of every CH row holds 16 there: 16 * 128 = 2048
    for (int k = tid; k < 3 * (kMmTR / 16); k += 32 * kMmEpiWarps)
      mm_sts128(a_mid + (k / (kMmTR / 16)) * kMmMidPlane + (k % (kMmTR / 16)) * (kMmKH * 16) + ((kMmKH - 1) >> 3) * 128 + ((kMmKH - 1) & 7) * 16,
                0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  mm_fence_before();
  __syncthreads();
  mm_fence_after();

  const uint32_t src_buf_bytes = 2u * L.R * 128u;
  const uint32_t cv_slot = (uint32_t)L.ksv_max * 2048u;
  // tile i lives in source buffer (and record slot) buf_of(i); it is that buffer's use number i / nbuf, whose parity the
  // buffer's barriers go by.  With ONE buffer the next tile's load waits for this tile's last reducev MMA: the tensor
  // pipe idles through a load, the epilogue warps (the bottleneck) mostly do not
  const bool two_bufs = L.nbuf == 2;
  auto buf_of = [&](int i) { return two_bufs ? (i & 1) : 0; };
  auto par_of = [&](int i) { return (uint32_t)((two_bufs ? (i >> 1) : i) & 1); };

  if (warp == kMmEpiWarps) {
    // =============================== producer ===============================
    if (lane == 0) {
      int i = 0, job = 0, chslot = 0, ch_uses[2] = {0, 0};
      const uint8_t* cv_in[4] = {nullptr, nullptr, nullptr, nullptr};
      const uint8_t* ch_in[2] = {nullptr, nullptr};
      for (;;) {
        const int s = atomicAdd(strip_counter, 1);
        if (s >= total_strips) break;
        while (job + 1 < n_jobs && s >= __ldg(&jobs[job + 1].strip_base)) job++;
        const MmJob* J = jobs + job;
        const int tiles_y = __ldg(&J->tiles_y), dh = __ldg(&J->dh), dw = __ldg(&J->dw), ksv = __ldg(&J->ksv);
        const unsigned long long pitch = __ldg(&J->dst_pitch);
        const int32_t* vfirst = mm_ldptr(&J->vfirst);
        const int32_t* hfirst = mm_ldptr(&J->hfirst);
        const int32_t* vmat = mm_ldptr(&J->vmat);
        const uint8_t* vmats = mm_ldptr(&J->vmats);
        const int strip = s - __ldg(&J->strip_base);
        const int ox0 = strip * kMmTC;
        const int bx0 = (3 * __ldg(hfirst + ox0)) & ~15;
        const uint8_t* hm = mm_ldptr(&J->hmats) + (size_t)__ldg(mm_ldptr(&J->hmat) + strip) * kMmChBytes;
        uint8_t* dst0 = const_cast<uint8_t*>(mm_ldptr(reinterpret_cast<const uint8_t* const*>(&J->dst))) + (size_t)__ldg(&J->dst_y0) * pitch +
                        (size_t)(__ldg(&J->dst_x0) + ox0) * 3;
        for (int rb = 0; rb < tiles_y; rb++, i++) {
          const int buf = buf_of(i);
          const int oy0 = rb * kMmTR;
          const int sy0 = __ldg(vfirst + oy0);
          int ws[4];
          const uint8_t* vm[4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            ws[q] = (__ldg(vfirst + min(oy0 + 32 * q, dh - 1)) - sy0) & ~7;
            vm[q] = vmats + (size_t)__ldg(vmat + 4 * rb + q) * ((size_t)ksv * 2048);
          }
          int flags = 0;
#pragma unroll
          for (int q = 0; q < 4; q++)
            if (vm[q] != cv_in[q]) flags |= 1 << q;
          if (hm != ch_in[chslot]) {
            chslot ^= 1;
            if (hm != ch_in[chslot]) flags |= 16;
          }
          flags |= chslot << 5;
          const unsigned long long dptr = (unsigned long long)(dst0 + (size_t)oy0 * pitch);
          twait(0, bar(kBarSrcFree + buf), par_of(i) ^ 1u);
          const uint32_t info = a_info + buf * (uint32_t)sizeof(MmInfo);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(info), "r"(dh - oy0), "r"(dw - ox0), "r"(bx0), "r"(1) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(info + 16), "r"(ws[0]), "r"(ws[1]), "r"(ws[2]), "r"(ws[3]) : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(info + 32), "r"(flags), "r"(ksv), "r"((uint32_t)dptr), "r"((uint32_t)(dptr >> 32)) : "memory");
          asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(info + 48), "r"((uint32_t)pitch), "r"((uint32_t)(pitch >> 32)) : "memory");
          const uint32_t src = a_src + (uint32_t)buf * src_buf_bytes;
          mbar_arrive_expect_tx(bar(kBarFull + buf), src_buf_bytes);
#pragma unroll
          for (int b = 0; b < 2; b++)
#pragma unroll
            for (int h = 0; h < 2; h++)
              tma_load_2d(src + b * (L.R * 128) + h * (L.R / 2) * 128, tmaps + job, bx0 + 128 * b, sy0 + h * (L.R / 2), bar(kBarFull + buf));
#pragma unroll
          for (int q = 0; q < 4; q++) {
            // the quarter's MMAs of the previous tile are done.  Waited for EVERY tile, reload or not: a parity wait
            // must never fall two phases behind its barrier
            twait(1, bar(kBarCvFree + q), (uint32_t)(i & 1) ^ 1u);
            if (flags & (1 << q)) {
              mbar_arrive_expect_tx(bar(kBarCvFull + q), (uint32_t)ksv * 2048u);
              mm_bulk_g2s(a_cv + q * cv_slot, vm[q], (uint32_t)ksv * 2048u, bar(kBarCvFull + q));
              cv_in[q] = vm[q];
            }
          }
          if (flags & 16) {
            twait(2, bar(kBarChFree + chslot), (uint32_t)(ch_uses[chslot] & 1) ^ 1u);
            mbar_arrive_expect_tx(bar(kBarChFull + chslot), kMmChBytes);
            mm_bulk_g2s(a_ch + chslot * kMmChBytes, hm, kMmChBytes, bar(kBarChFull + chslot));
            ch_in[chslot] = hm;
          }
          ch_uses[chslot]++;
        }
      }
      // the record that ends the sequence
      const int buf = buf_of(i);
      mbar_wait(bar(kBarSrcFree + buf), par_of(i) ^ 1u);
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(a_info + buf * (uint32_t)sizeof(MmInfo) + 12), "r"(0) : "memory");
      mbar_arrive(bar(kBarFull + buf));
    }
  } else if (warp > kMmEpiWarps) {
    // =============================== MMA issue ===============================
    // Three issuing warps (one lane each): w = column block of reducev (0, 1) and channel plane of reduceh (0, 1, 2).
    // Measured (tools/probes/umma_rate.cu): two or more WARPS issuing MMAs with a commit every few of them keep the
    // tensor pipe at its floor for this shape (51.5 cycles per M128 N64 K32); lanes of one warp do not (57-93), and a
    // lone thread that never commits gets 125; commits are nearly free there.  So each warp commits to everything that
    // waits for its MMAs: accumulator complete, matrix slot free, source buffer free.  (Only the CH slot is relayed
    // by warp 0 with a plain arrive when the next reduceh is issued: reduceh is serialised by its one accumulator.)
    const int w = warp - kMmEpiWarps - 1;
    uint32_t vuse0 = 0, vuse1 = 0, hcount = 0;
    uint32_t cvl0 = 0, cvl1 = 0, cvl2 = 0, cvl3 = 0, chl0 = 0, chl1 = 0;
    int hprev_slot = 0;
    auto issue_v = [&](const MmInfo& inf, int buf, int q) {   // q is a literal at every call site
      const int a = q & 1;
      uint32_t& vuse = a ? vuse1 : vuse0;
      uint32_t& cvl = q == 0 ? cvl0 : (q == 1 ? cvl1 : (q == 2 ? cvl2 : cvl3));
      twait(0, bar(kBarVFree + a), (vuse & 1u) ^ 1u);
      vuse++;
      if (inf.flags & (1 << q)) {
        twait(1, bar(kBarCvFull + q), cvl & 1u);
        cvl++;
      }
      mm_fence_after();
      const int wsq = q == 0 ? inf.ws0 : (q == 1 ? inf.ws1 : (q == 2 ? inf.ws2 : inf.ws3));
      const uint32_t src = a_src + (uint32_t)buf * src_buf_bytes + w * (L.R * 128) + wsq * 128;
      for (int ks = 0; ks < inf.ksv; ks++)   // the first K step overwrites the accumulator: nothing to clear or re-arm
        mm_mma_i8(tm + kMmColV + a * 128 + w * 64, mm_desc(src + ks * 4096, (uint32_t)L.R * 128u, 1024u, 2u),
                  mm_desc(a_cv + q * cv_slot + ks * 2048, 1024u, 128u, 0u), kMmIdesc, ks > 0);
      mm_commit(bar(kBarVFull + a));
      mm_commit(bar(kBarCvFree + q));                  // the quarter's matrix slot can be refilled
      if (q == 3) mm_commit(bar(kBarSrcFree + buf));   // ... and the tile's source buffer
    };
    auto issue_h = [&](const MmInfo& inf) {
      twait(2, bar(kBarMidFull), hcount & 1u);
      twait(3, bar(kBarHFree), (hcount & 1u) ^ 1u);
      if (w == 0 && hcount > 0) mbar_arrive(bar(kBarChFree + hprev_slot));
      hcount++;
      const int slot = (inf.flags >> 5) & 1;
      hprev_slot = slot;
      if (inf.flags & 16) {
        uint32_t& chl = slot ? chl1 : chl0;
        twait(4, bar(kBarChFull + slot), chl & 1u);
        chl++;
      }
      mm_fence_after();
#pragma unroll
      for (int ks = 0; ks < kMmKH / 32; ks++)
        mm_mma_i8(tm + kMmColH + 64 * w, mm_desc(a_mid + w * kMmMidPlane + ks * 512, 128u, (uint32_t)kMmKH * 16u, 0u),
                  mm_desc(a_ch + slot * kMmChBytes + ks * 2048, 1024u, 128u, 0u), kMmIdesc, ks > 0);
      mm_commit(bar(kBarHFull));
    };
    if (lane == 0) {
    mbar_wait(bar(kBarFull), 0);
    MmInfo cur = mm_load_info(a_info), prev = cur;
    if (cur.valid) {
      if (w < 2) issue_v(cur, 0, 0);
      for (int i = 0;; i++) {
        const int buf = buf_of(i), nbuf = buf_of(i + 1);
        if (i > 0) issue_h(prev);
        if (w < 2) {
          issue_v(cur, buf, 1);
          issue_v(cur, buf, 2);
          issue_v(cur, buf, 3);
        }
        twait(5, bar(kBarFull + nbuf), par_of(i + 1));
        const MmInfo nxt = mm_load_info(a_info + nbuf * (uint32_t)sizeof(MmInfo));
        if (nxt.valid && w < 2) issue_v(nxt, nbuf, 0);
        prev = cur;
        cur = nxt;
        if (!cur.valid) {
          issue_h(prev);
          break;
        }
      }
    }
    }
  } else {
    // =============================== epilogue warps ===============================
    uint32_t hcnt = 0;
    mbar_wait(bar(kBarFull), 0);
    MmInfo cur = mm_load_info(a_info), prev = cur;
    // reduceh epilogue of tile `inf`: thread = output row, 8 pixels x RGB -> 24 interleaved bytes into the staging tile;
    // then the 16 warps write the tile out row by row (a warp instruction covers 2.7 rows of 96 contiguous bytes
    // instead of 32 rows of 8).  No barrier guards the staging tile against the NEXT tile's writes: between this
    // read and that write every warp passes two accumulator hand-overs that wait for all 16 warps.
    auto epi_h = [&](const MmInfo& inf) {
      twait(2, bar(kBarHFull), hcnt & 1u);
      hcnt++;
      mm_fence_after();
      if (prof) tk = mm_clock();
      uint32_t hi[3][8], lo[3][8];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        mm_ld8(tlane + kMmColH + 64 * c + 8 * pg, hi[c]);
        mm_ld8(tlane + kMmColH + 64 * c + 32 + 8 * pg, lo[c]);
      }
      mm_wait_ld();
      mm_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBarHFree));   // the accumulators are in registers: the next reduceh may overwrite them
      int v[24];   // interleaved: v[3 j + c]
#pragma unroll
      for (int j = 0; j < 8; j++)
#pragma unroll
        for (int c = 0; c < 3; c++) v[3 * j + c] = ((int)hi[c][j] * 128 + (int)lo[c][j]) >> IRP_INTERP_SHIFT;
      const uint32_t so = a_out + (32 * lq + lane) * kMmOutPitch + 24 * pg;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const uint32_t w0 = pack_sat_u8(v[8 * k + 1], v[8 * k], pack_sat_u8(v[8 * k + 3], v[8 * k + 2], 0u));
        const uint32_t w1 = pack_sat_u8(v[8 * k + 5], v[8 * k + 4], pack_sat_u8(v[8 * k + 7], v[8 * k + 6], 0u));
        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(so + 8 * k), "r"(w0), "r"(w1) : "memory");
      }
      tick(10);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kMmEpiWarps) : "memory");
      uint8_t* dst = reinterpret_cast<uint8_t*>(((unsigned long long)inf.dst_hi << 32) | inf.dst_lo);
      const int nbytes = 3 * min(inf.cols, kMmTC);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int idx = tid + 32 * kMmEpiWarps * k, row = idx / 12, c8 = 8 * (idx - 12 * row);
        if (row < inf.rows && c8 < nbytes) {
          uint32_t w0, w1;
          asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(a_out + row * kMmOutPitch + c8));
          uint8_t* d = dst + (size_t)row * inf.pitch + c8;
          if (c8 + 8 <= nbytes) {
            *reinterpret_cast<uint2*>(d) = make_uint2(w0, w1);
          } else {   // the image's last columns
            const unsigned long long ww = ((unsigned long long)w1 << 32) | w0;
            for (int j = 0; j < nbytes - c8; j++) d[j] = (uint8_t)(ww >> (8 * j));
          }
        }
      }
      tick(11);
    };
    if (cur.valid) {
      for (int i = 0;; i++) {
        const int nbuf = buf_of(i + 1);
        // planar address of this thread's source byte column (reducev epilogue)
        uint32_t mid_col;
        {
          const int xb = 128 * vb + 32 * lq + lane, B = cur.bx0 + xb, p = B / 3, c = B - 3 * p, kpx = p - cur.bx0 / 3;
          mid_col = a_mid + c * kMmMidPlane + (kpx >> 3) * 128 + (kpx & 7) * 16;
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int a = q & 1;
          twait(0, bar(kBarVFull + a), (uint32_t)(q >> 1));   // each accumulator completes twice per tile: phases 2 i, 2 i + 1
          mm_fence_after();
          if (prof) tk = mm_clock();
          uint32_t hi[16], lo[16];
          const uint32_t ta = tlane + kMmColV + a * 128 + vb * 64 + 16 * vrh;
          mm_ld16(ta, hi);
          mm_ld16(ta + 32, lo);
          mm_wait_ld();
          tick(6);
          mm_fence_before();   // hand the accumulator back before anything else: the next quarter's first MMA overwrites it
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(kBarVFree + a));
          tick(7);
          const uint32_t p0 = mm_pack4(hi, lo), p1 = mm_pack4(hi + 4, lo + 4), p2 = mm_pack4(hi + 8, lo + 8), p3 = mm_pack4(hi + 12, lo + 12);
          tick(8);
          if (q == 0 && i > 0) {
            // the previous tile's reduceh: its epilogue first (that wait also means the intermediate has been read),
            // then this quarter's bytes may go in
            epi_h(prev);
            if (prof) tk = mm_clock();
          }
          mm_sts128(mid_col + (2 * q + vrh) * (kMmKH * 16), p0, p1, p2, p3);
          if (q == 3) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the intermediate is read by the tensor core next
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kBarMidFull));
          }
          tick(9);
        }
        twait(3, bar(kBarFull + nbuf), par_of(i + 1));
        const MmInfo nxt = mm_load_info(a_info + nbuf * (uint32_t)sizeof(MmInfo));
        prev = cur;
        cur = nxt;
        if (!cur.valid) {
          epi_h(prev);
          break;
        }
      }
    }
  }
  if (prof && lane == 0 && (warp == 0 || warp >= kMmEpiWarps)) {   // rows: epilogue warp 0, producer, issuing warps
    const int r = warp == 0 ? 0 : warp - kMmEpiWarps + 1;
    for (int k = 0; k < 12; k++) dbg[16 * r + k] = wacc[k];
    dbg[16 * r + 15] = mm_clock() - t_start;
  }
  __syncwarp();
  mm_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

}  // namespace irp
