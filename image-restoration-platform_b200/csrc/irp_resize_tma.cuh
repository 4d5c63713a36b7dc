// irp_resize_tma.cuh — the streaming variant of the lanczos3 resize kernel for the common case
// (3-channel sources with 16-byte aligned base and pitch, a real shrink: more than one tap).
//
// Same arithmetic as resize_kernel<3> (irp_resize.cuh; reference imagePreprocess.js:46-53 = libvips
// reducev then reduceh, 12-bit fixed point, u8 between the passes), different data movement:
//   * one CTA per SM, four or five independent 4-warp groups, each on its own output tile;
//   * the tile's source footprint arrives by ONE 2-D TMA tile load (cp.async.bulk.tensor, the image
//     described as rows of u16), its per-row vertical table and per-column horizontal table by two
//     plain bulk copies — all issued by one lane one tile ahead and signalled on mbarriers, so no
//     warp ever waits on HBM or on a coefficient lookup;
//   * the raw rows are turned into the row-pair-interleaved layout IDP.2A wants IN PLACE (each warp
//     reads a row pair into registers, then writes it back interleaved), still channel-interleaved:
//     the vertical pass does not care about channels;
//   * reducev works on 12-byte (4-pixel) columns and de-interleaves its u8 result on the way to the
//     planar intermediate tile (half the bytes of the source, so half the PRMT work);
//   * reduceh is the planar code of irp_resize.cuh with its coefficients read from shared memory.
// Bytes outside the image arrive as zeros and are replaced by replicate rows / columns first.
#pragma once
#include "irp_classify_bulk.cuh"
#include "irp_resize.cuh"

namespace irp {

// 4-warp groups per CTA and rows per output tile are chosen per launch: 5 groups x 24-row tiles when every
// job's footprint fits a fifth of the shared memory (12 MP -> 2048: 2 % faster than 4 x 32; 6 x 16 is 6 %
// slower), else 4 groups x 32-row tiles (larger shrinks keep full-width tiles).
constexpr int kRMaxGroups = 5;
constexpr int kRTow = 64, kRTohMax = 32;             // largest output tile
constexpr int kTabWords = 16;                        // words per row of the vertical table
constexpr int kHTabWords = 20;                       // ... of the horizontal table: 80-byte rows keep the per-lane LDS.128 conflict-free

struct RtJob {                   // one image of a streaming resize launch
  uint8_t* dst;
  unsigned long long dst_pitch;
  const uint32_t* vrows;         // [dh][16]: 13 vertical coefficient pairs (even-row aligned), word 15 = first pair-row
  const uint32_t* hcols;         // [dw][20]: 14 horizontal coefficient pairs (word aligned),   word 15 = first word column
  const int32_t* vstart;         // [dh] first tap row (unclamped)
  const int32_t* hstart;         // [dw] first tap column (unclamped)
  int sw, sh, dw, dh;
  int dst_x0, dst_y0;
  int tow, toh, tiles_x, tiles_y, tile_base;
  int vn, hn;
  int box_cols, box_rows;        // this job's TMA box: bytes per row (multiple of 16), rows (even); <= the launch-wide layout
  int pad[1];
};

struct RtLayout {                // byte offsets inside one group's shared-memory slice (uniform per launch)
  int box_cols, box_rows;        // largest TMA box of the launch: bytes per row (multiple of 16), rows (even)
  int mid_pitch;                 // bytes per row per plane of the intermediate tile
  int off_mid, off_vtab, off_hcols, off_sync, group_bytes;
  int groups, mid_rows;          // groups per CTA (blockDim.x / 128), rows reserved per plane of the intermediate tile
};

struct RtInfo {                  // next tile, written by the issuing lane
  int job, ox0, oy0, ow;
  int oh, sx0, sy0, ntr;         // oh = 0: end of the sequence; sx0 (multiple of 4) / sy0 (even): image coordinates of the
};                               // box origin; ntr: 4-pixel source columns the tile needs

// reducev for one 12-byte column and output rows (r, r + 1) whose windows start D pair-rows apart.
// sp: the column in the pair-interleaved tile at the first row's window; results de-interleaved into
// the planar intermediate tile.  TWO = false computes row r only.
template <int NPV, int D, bool TWO>
__device__ __forceinline__ void vpass_triple(uint32_t sp, int pair_pitch, const uint32_t* vt0, uint32_t mid_addr, int mid_pitch,
                                             int mid_plane) {
  uint32_t c0[16], c1[16];
#pragma unroll
  for (int q4 = 0; q4 < (NPV + 3) / 4; q4++) {
    const uint4 a = reinterpret_cast<const uint4*>(vt0)[q4];
    c0[4 * q4] = a.x; c0[4 * q4 + 1] = a.y; c0[4 * q4 + 2] = a.z; c0[4 * q4 + 3] = a.w;
    if (TWO) {
      const uint4 b = reinterpret_cast<const uint4*>(vt0 + kTabWords)[q4];
      c1[4 * q4] = b.x; c1[4 * q4 + 1] = b.y; c1[4 * q4 + 2] = b.z; c1[4 * q4 + 3] = b.w;
    }
  }
  int a[12], b[12];
#pragma unroll
  for (int j = 0; j < 12; j++) a[j] = b[j] = 1 << (IRP_INTERP_SHIFT - 1);
#pragma unroll
  for (int pp = 0; pp < NPV + (TWO ? D : 0); pp++) {
    uint32_t w[6];
#pragma unroll
    for (int k = 0; k < 3; k++)
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w[2 * k]), "=r"(w[2 * k + 1]) : "r"(sp + pp * pair_pitch + 8 * k));
    if (pp < NPV) {
#pragma unroll
      for (int k = 0; k < 6; k++) {
        a[2 * k] = dp2a_lo_s16_u8(c0[pp], w[k], a[2 * k]);
        a[2 * k + 1] = dp2a_hi_s16_u8(c0[pp], w[k], a[2 * k + 1]);
      }
    }
    if (TWO && pp >= D) {
#pragma unroll
      for (int k = 0; k < 6; k++) {
        b[2 * k] = dp2a_lo_s16_u8(c1[pp - D], w[k], b[2 * k]);
        b[2 * k + 1] = dp2a_hi_s16_u8(c1[pp - D], w[k], b[2 * k + 1]);
      }
    }
  }
#pragma unroll
  for (int row = 0; row < (TWO ? 2 : 1); row++) {
    const int* v = row ? b : a;
    uint32_t w[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const uint32_t hi = pack_sat_u8(v[4 * k + 3] >> IRP_INTERP_SHIFT, v[4 * k + 2] >> IRP_INTERP_SHIFT, 0u);
      w[k] = pack_sat_u8(v[4 * k + 1] >> IRP_INTERP_SHIFT, v[4 * k] >> IRP_INTERP_SHIFT, hi);
    }
    // w0 = R0 G0 B0 R1 | w1 = G1 B1 R2 G2 | w2 = B2 R3 G3 B3
    const uint32_t R = __byte_perm(__byte_perm(w[0], w[1], 0x0630), w[2], 0x5210);
    const uint32_t G = __byte_perm(__byte_perm(w[0], w[1], 0x0741), w[2], 0x6210);
    const uint32_t B = __byte_perm(__byte_perm(w[0], w[1], 0x0052), w[2], 0x7410);
    const uint32_t m = mid_addr + row * mid_pitch;
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(m), "r"(R) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(m + mid_plane), "r"(G) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(m + 2 * mid_plane), "r"(B) : "memory");
  }
}

template <int NPV>
__device__ __forceinline__ void vpass_item(uint32_t sp, int pair_pitch, const uint32_t* vt0, bool two, int sy0_pair, uint32_t mid_addr,
                                           int mid_pitch, int mid_plane) {
  const int q0 = (int)vt0[15] - sy0_pair;
  const uint32_t p = sp + q0 * pair_pitch;
  if (two) {
    const int d = (int)vt0[kTabWords + 15] - (int)vt0[15];   // uniform over the group
    if (d == 1)
      vpass_triple<NPV, 1, true>(p, pair_pitch, vt0, mid_addr, mid_pitch, mid_plane);
    else if (d == 0)
      vpass_triple<NPV, 0, true>(p, pair_pitch, vt0, mid_addr, mid_pitch, mid_plane);
    else if (d == 2)
      vpass_triple<NPV, 2, true>(p, pair_pitch, vt0, mid_addr, mid_pitch, mid_plane);
    else {
      vpass_triple<NPV, 0, false>(p, pair_pitch, vt0, mid_addr, mid_pitch, mid_plane);
      vpass_triple<NPV, 0, false>(p + d * pair_pitch, pair_pitch, vt0 + kTabWords, mid_addr + mid_pitch, mid_pitch, mid_plane);
    }
  } else {
    vpass_triple<NPV, 0, false>(p, pair_pitch, vt0, mid_addr, mid_pitch, mid_plane);
  }
}

// reduceh for one output column over rows [r_begin, r_end), three planes; coefficient pairs from shared memory
template <int NWH>
__device__ __forceinline__ void hpass_column_s(uint32_t mbase, int mid_pitch, int mid_plane, const uint32_t* hp, uint8_t* d,
                                               unsigned long long dst_pitch, int r_begin, int r_end) {
  uint32_t cph[2 * NWH + 2];
#pragma unroll
  for (int q4 = 0; q4 < (2 * NWH + 3) / 4; q4++) {
    const uint4 c4 = reinterpret_cast<const uint4*>(hp)[q4];
    if (4 * q4 < 2 * NWH + 2) cph[4 * q4] = c4.x;
    if (4 * q4 + 1 < 2 * NWH + 2) cph[4 * q4 + 1] = c4.y;
    if (4 * q4 + 2 < 2 * NWH + 2) cph[4 * q4 + 2] = c4.z;
    if (4 * q4 + 3 < 2 * NWH + 2) cph[4 * q4 + 3] = c4.w;
  }
  // two rows per iteration: six independent IDP chains hide the dot-product latency
  int r = r_begin;
  for (; r + 1 < r_end; r += 2) {
    uint32_t v[2][3];
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
#pragma unroll
      for (int ch = 0; ch < 3; ch++) {
        const uint32_t ma = mbase + ch * mid_plane + (r + rr) * mid_pitch;
        uint32_t w[NWH];
#pragma unroll
        for (int wv = 0; wv < NWH; wv++) w[wv] = lds_b32(ma + 4 * wv);
        int acc = 1 << (IRP_INTERP_SHIFT - 1);
#pragma unroll
        for (int wv = 0; wv < NWH; wv++) {
          acc = dp2a_lo_s16_u8(cph[2 * wv], w[wv], acc);
          acc = dp2a_hi_s16_u8(cph[2 * wv + 1], w[wv], acc);
        }
        v[rr][ch] = pack_sat_u8(0, acc >> IRP_INTERP_SHIFT, 0u);
      }
    }
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
      uint8_t* dp = d + (size_t)(r + rr) * dst_pitch;
      dp[0] = (uint8_t)v[rr][0];
      dp[1] = (uint8_t)v[rr][1];
      dp[2] = (uint8_t)v[rr][2];
    }
  }
  if (r < r_end) {
    uint32_t v[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      const uint32_t ma = mbase + ch * mid_plane + r * mid_pitch;
      uint32_t w[NWH];
#pragma unroll
      for (int wv = 0; wv < NWH; wv++) w[wv] = lds_b32(ma + 4 * wv);
      int acc = 1 << (IRP_INTERP_SHIFT - 1);
#pragma unroll
      for (int wv = 0; wv < NWH; wv++) {
        acc = dp2a_lo_s16_u8(cph[2 * wv], w[wv], acc);
        acc = dp2a_hi_s16_u8(cph[2 * wv + 1], w[wv], acc);
      }
      v[ch] = pack_sat_u8(0, acc >> IRP_INTERP_SHIFT, 0u);
    }
    uint8_t* dp = d + (size_t)r * dst_pitch;
    dp[0] = (uint8_t)v[0];
    dp[1] = (uint8_t)v[1];
    dp[2] = (uint8_t)v[2];
  }
}

struct RtIssuer {
  int next_tile, stride, job;
  const uint32_t* hsrc;   // the pending tile's slice of the horizontal table
  int hbytes;
};

// part A (the source tile and the vertical table are free): describe the group's next tile, start its
// TMA tile load and the bulk copy of its vertical-table rows
__device__ __forceinline__ void rt_issue_tile(RtIssuer& is, const RtJob* __restrict__ jobs, const TmaDesc* __restrict__ tmaps, int n_jobs,
                                              int total_tiles, const RtLayout& L, uint32_t bar_tile, uint32_t info_addr, uint32_t a_src,
                                              uint32_t a_vtab) {
  const int tile = is.next_tile;
  is.next_tile += is.stride;
  is.hbytes = 0;
  if (tile >= total_tiles) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(info_addr + 16), "r"(0) : "memory");
    mbar_arrive(bar_tile);
    return;
  }
  while (is.job + 1 < n_jobs && tile >= __ldg(&jobs[is.job + 1].tile_base)) is.job++;
  const RtJob* J = jobs + is.job;
  const int tiles_x = __ldg(&J->tiles_x), tow = __ldg(&J->tow), toh = __ldg(&J->toh), dw = __ldg(&J->dw), dh = __ldg(&J->dh);
  const int t = tile - __ldg(&J->tile_base);
  const int ty = t / tiles_x, tx = t - ty * tiles_x;
  const int ox0 = tx * tow, oy0 = ty * toh;
  const int ow = min(tow, dw - ox0), oh = min(toh, dh - oy0);
  const int32_t* vstart = reinterpret_cast<const int32_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&J->vstart)));
  const int32_t* hstart = reinterpret_cast<const int32_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&J->hstart)));
  const uint32_t* vrows = reinterpret_cast<const uint32_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&J->vrows)));
  const uint32_t* hcols = reinterpret_cast<const uint32_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&J->hcols)));
  const int sy0 = __ldg(vstart + oy0) & ~1, sx0 = __ldg(hstart + ox0) & ~3;
  const int ntr = (__ldg(hstart + ox0 + ow - 1) + __ldg(&J->hn) - sx0 + 3) >> 2;
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(info_addr), "r"(is.job), "r"(ox0), "r"(oy0), "r"(ow) : "memory");
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(info_addr + 16), "r"(oh), "r"(sx0), "r"(sy0), "r"(ntr) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes (in-place interleave) before the async-proxy refill
  mbar_arrive_expect_tx(bar_tile, (uint32_t)(__ldg(&J->box_rows) * __ldg(&J->box_cols) + oh * kTabWords * 4));
  tma_load_2d(a_src, tmaps + is.job, ((sx0 * 3) & ~15) >> 1, sy0, bar_tile);   // a box row must start on a 16-byte boundary
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(a_vtab),
               "l"(vrows + (size_t)oy0 * kTabWords), "r"(oh * kTabWords * 4), "r"(bar_tile)
               : "memory");
  is.hsrc = hcols + (size_t)ox0 * kHTabWords;
  is.hbytes = ow * kHTabWords * 4;
}
// part B (the horizontal table is free too): the pending tile's horizontal-table columns
__device__ __forceinline__ void rt_issue_hcols(const RtIssuer& is, uint32_t bar_h, uint32_t a_hcols) {
  if (!is.hbytes) return;
  mbar_arrive_expect_tx(bar_h, (uint32_t)is.hbytes);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(a_hcols), "l"(is.hsrc),
               "r"(is.hbytes), "r"(bar_h)
               : "memory");
}

__global__ void __launch_bounds__(kRMaxGroups * 128, 1)
resize_tma_kernel(const RtJob* __restrict__ jobs, const TmaDesc* __restrict__ tmaps, int n_jobs, int total_tiles, RtLayout L) {
  extern __shared__ __align__(128) uint8_t rsmem[];
  __shared__ __align__(16) RtJob s_jobs[kRMaxGroups];
  const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(rsmem) + 127u) & ~127u;
  const int group = threadIdx.x >> 7, tid = threadIdx.x & 127, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_src = sbase + group * L.group_bytes;
  const uint32_t a_mid = a_src + L.off_mid, a_vtab = a_src + L.off_vtab, a_hcols = a_src + L.off_hcols;
  const uint32_t a_bar = a_src + L.off_sync, a_barh = a_bar + 8, a_info = a_bar + 16;
  // generic views of the two tables (read with ordinary loads)
  uint8_t* g_base = rsmem + (a_src - (uint32_t)__cvta_generic_to_shared(rsmem));
  const uint32_t* s_vtab = reinterpret_cast<const uint32_t*>(g_base + L.off_vtab);
  const uint32_t* s_hcols = reinterpret_cast<const uint32_t*>(g_base + L.off_hcols);
  RtJob& SJ = s_jobs[group];
  if (tid == 0) {
    mbar_init(a_bar, 1);
    mbar_init(a_barh, 1);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  RtIssuer is;
  is.next_tile = blockIdx.x * L.groups + group;
  is.stride = gridDim.x * L.groups;
  is.job = 0;
  is.hsrc = nullptr;
  is.hbytes = 0;
  if (tid == 0) {
    rt_issue_tile(is, jobs, tmaps, n_jobs, total_tiles, L, a_bar, a_info, a_src, a_vtab);
    rt_issue_hcols(is, a_barh, a_hcols);
  }

  const int mid_plane = L.mid_rows * L.mid_pitch;
  int loaded_job = -1;
  for (uint32_t it = 0;; it++) {
    mbar_wait(a_bar, it & 1u);
    const uint32_t ia = a_info + (it & 1u) * 32u;
    int job, ox0, oy0, ow, oh, sx0, sy0, ntr;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(job), "=r"(ox0), "=r"(oy0), "=r"(ow) : "r"(ia));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(oh), "=r"(sx0), "=r"(sy0), "=r"(ntr) : "r"(ia + 16));
    if (oh == 0) break;
    if (job != loaded_job) {   // uniform over the group
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>(jobs + job);
      for (int k = tid; k < (int)(sizeof(RtJob) / 4); k += 128) reinterpret_cast<uint32_t*>(&SJ)[k] = __ldg(s32 + k);
      loaded_job = job;
      group_barrier(group);
    }
    const int sw = SJ.sw, sh = SJ.sh;
    const int box_cols = SJ.box_cols, box_rows = SJ.box_rows, pair_pitch = 2 * box_cols;
    const int delta = (sx0 * 3) & 15;   // byte offset of pixel sx0 inside a box row (0, 4, 8 or 12)

    // ---- replicate rows / columns that fell outside the image (TMA wrote zeros there) ----
    if (sy0 < 0 || sy0 + box_rows > sh) {
      const int lo = max(-sy0, 0), hi = min(sh - 1 - sy0, box_rows - 1);   // box rows holding image rows 0 and sh - 1
      for (int r = warp; r < box_rows; r += 4) {
        const int sr = min(max(r, lo), hi);
        if (sr != r)
          for (int k = lane; k < box_cols / 4; k += 32) {
            const uint32_t v = lds_b32(a_src + sr * box_cols + 4 * k);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(a_src + r * box_cols + 4 * k), "r"(v) : "memory");
          }
      }
      group_barrier(group);
    }
    if (sx0 < 0 || sx0 + (box_cols - delta) / 3 > sw) {
      const int ncol = (box_cols - delta) / 3;
      const int lo = max(-sx0, 0), hi = min(sw - 1 - sx0, ncol - 1);          // box columns holding image columns 0 and sw - 1
      for (int r = tid; r < box_rows; r += 128) {
        const uint32_t rr = a_src + r * box_cols + delta;
        for (int j = 0; j < lo; j++)
          for (int k = 0; k < 3; k++) sts_u8(rr + 3 * j + k, lds_u8(rr + 3 * lo + k));
        for (int j = hi + 1; j < ncol; j++)
          for (int k = 0; k < 3; k++) sts_u8(rr + 3 * j + k, lds_u8(rr + 3 * hi + k));
      }
      group_barrier(group);
    }

    // ---- rows (2q, 2q+1) -> pair-interleaved words (a_k, b_k, a_k+1, b_k+1), in place, one warp per pair-row ----
    for (int q = warp; q < box_rows / 2; q += 4) {
      const uint32_t ra = a_src + q * pair_pitch;
      uint4 va = make_uint4(0, 0, 0, 0), vb = va;
      const bool on = lane * 16 < box_cols;
      if (on) {
        va = lds_v4(ra + 16 * lane);
        vb = lds_v4(ra + box_cols + 16 * lane);
      }
      __syncwarp();
      if (on) {
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ra + 32 * lane), "r"(__byte_perm(va.x, vb.x, 0x5140)),
                     "r"(__byte_perm(va.x, vb.x, 0x7362)), "r"(__byte_perm(va.y, vb.y, 0x5140)), "r"(__byte_perm(va.y, vb.y, 0x7362))
                     : "memory");
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ra + 32 * lane + 16), "r"(__byte_perm(va.z, vb.z, 0x5140)),
                     "r"(__byte_perm(va.z, vb.z, 0x7362)), "r"(__byte_perm(va.w, vb.w, 0x5140)), "r"(__byte_perm(va.w, vb.w, 0x7362))
                     : "memory");
      }
    }
    group_barrier(group);

    // ---- reducev: (row pair, 12-byte column) items -> planar u8 intermediate tile ----
    {
      const int npv = (SJ.vn + 2) >> 1;
      const int nrp = (oh + 1) >> 1;
      const int sy0_pair = sy0 >> 1;
      // A warp takes one row pair at a time and its lanes the first 32 columns, so the window distance
      // d is uniform in the warp and the LDS.64 stream is conflict-free; columns beyond 32 (a 64-wide
      // tile at shrink 1.95 has 35) are left-over items handled afterwards.
      const int nmain = min(ntr, 32), nleft = ntr - nmain;
      for (int rp = warp; rp < nrp; rp += 4) {
        if (lane >= nmain) continue;
        const int r = 2 * rp;
        const bool two = r + 1 < oh;
        const uint32_t sp = a_src + 2 * delta + lane * 24;
        const uint32_t* vt0 = s_vtab + r * kTabWords;
        const uint32_t ma = a_mid + r * L.mid_pitch + lane * 4;
        switch (npv) {
          case 2: vpass_item<2>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
          case 3: vpass_item<3>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
          case 4: vpass_item<4>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
          case 5: vpass_item<5>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
          case 6: vpass_item<6>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
          case 7: vpass_item<7>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
          case 8: vpass_item<8>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
          case 9: vpass_item<9>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
          case 10: vpass_item<10>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
          case 11: vpass_item<11>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
          case 12: vpass_item<12>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
          default: vpass_item<13>(sp, pair_pitch, vt0, two, sy0_pair, ma, L.mid_pitch, mid_plane); break;
        }
      }
      // left-over columns, one output ROW per item so that they spread over the whole group
      for (int j = tid; j < nleft * oh; j += 128) {
        const int r = j / nleft, tr = 32 + j - r * nleft;
        const uint32_t* vt0 = s_vtab + r * kTabWords;
        const uint32_t sp = a_src + 2 * delta + tr * 24 + ((int)vt0[15] - sy0_pair) * pair_pitch;
        const uint32_t ma = a_mid + r * L.mid_pitch + tr * 4;
        switch (npv) {
          case 2: vpass_triple<2, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
          case 3: vpass_triple<3, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
          case 4: vpass_triple<4, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
          case 5: vpass_triple<5, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
          case 6: vpass_triple<6, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
          case 7: vpass_triple<7, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
          case 8: vpass_triple<8, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
          case 9: vpass_triple<9, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
          case 10: vpass_triple<10, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
          case 11: vpass_triple<11, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
          case 12: vpass_triple<12, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
          default: vpass_triple<13, 0, false>(sp, pair_pitch, vt0, ma, L.mid_pitch, mid_plane); break;
        }
      }
    }
    group_barrier(group);

    // ---- the source tile and the vertical table are free: start this group's next tile ----
    if (tid == 0) rt_issue_tile(is, jobs, tmaps, n_jobs, total_tiles, L, a_bar, a_info + ((it + 1) & 1u) * 32u, a_src, a_vtab);

    // ---- reduceh + store ----
    mbar_wait(a_barh, it & 1u);   // this tile's horizontal table (landed long ago)
    {
      const int xcol = tid & 63, rgh = tid >> 6;
      const int rows_per = (oh + 1) >> 1;
      const int r_begin = rgh * rows_per, r_end = min(oh, r_begin + rows_per);
      const uint32_t* hp = s_hcols + xcol * kHTabWords;
      const int hn = SJ.hn;
      const int nwh = (hn + 3 + 3) >> 2;
      uint8_t* d = SJ.dst + (size_t)(SJ.dst_y0 + oy0) * SJ.dst_pitch + (size_t)(SJ.dst_x0 + ox0 + xcol) * 3;
      const unsigned long long dpitch = SJ.dst_pitch;
      if (xcol < ow) {
        const uint32_t mbase = a_mid + ((int)hp[15] - (sx0 >> 2)) * 4;
        switch (nwh) {
          case 2: hpass_column_s<2>(mbase, L.mid_pitch, mid_plane, hp, d, dpitch, r_begin, r_end); break;
          case 3: hpass_column_s<3>(mbase, L.mid_pitch, mid_plane, hp, d, dpitch, r_begin, r_end); break;
          case 4: hpass_column_s<4>(mbase, L.mid_pitch, mid_plane, hp, d, dpitch, r_begin, r_end); break;
          case 5: hpass_column_s<5>(mbase, L.mid_pitch, mid_plane, hp, d, dpitch, r_begin, r_end); break;
          case 6: hpass_column_s<6>(mbase, L.mid_pitch, mid_plane, hp, d, dpitch, r_begin, r_end); break;
          default: hpass_column_s<7>(mbase, L.mid_pitch, mid_plane, hp, d, dpitch, r_begin, r_end); break;
        }
      }
    }
    group_barrier(group);
    if (tid == 0) rt_issue_hcols(is, a_barh, a_hcols);
  }
}

}  // namespace irp
