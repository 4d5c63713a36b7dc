// irp_jpeg_prog.cuh — multi-scan JPEG on the device (sm_100a): progressive files (SOF2: spectral selection +
// successive approximation — what `sharp(...).jpeg({progressive: true, mozjpeg: true})` writes,
// server-node/src/middleware/imagePreprocess.js:57-61, and therefore what the classifier is handed in the planned
// worker flow; accepted as uploads by uploadValidation.js:7) and sequential files with one scan per component.
// The scans fill the SAME coefficient arena and compact DC array the baseline Huffman stage fills
// (irp_jpeg.cuh: zigzag order, 64 int16 per block), so dequantisation, IDCT, upsampling and colour conversion are
// shared and the pixels are libjpeg-turbo's (its jdphuff.c is the algorithm; the tests hold the result against Pillow).
//
// Parallelism.  A progressive scan is one Huffman stream whose state (band position, end-of-band run, and for the
// refinements the history of every coefficient) does not re-synchronise the way a baseline stream does, so a scan is
// decoded by ONE warp: lane 0 walks the symbols, the other lanes are its memory system —
//   * the stream reaches lane 0 through a shared-memory ring the warp refills with coalesced loads (the bit reader
//     never waits on global memory),
//   * a refinement scan needs, per block, which coefficients are already nonzero and has to update them in place:
//     the warp loads eight blocks at a time (lane = zigzag positions l and l + 32: two coalesced 64-byte rows per
//     block), ballots give lane 0 a 64-bit history mask per block, lane 0 turns symbols and correction bits into
//     three masks per block (correction, new coefficient, its sign) and the warp applies them and stores,
//   * a DC refinement is one bit per block at a known position: every lane takes its own blocks.
// The scans of a batch are independent across images and across the coefficient sets they touch.  A scan that touches
// coefficients an earlier scan of the same image wrote (a refinement after its first pass) does NOT wait for that
// scan to end: both walk the component's blocks in the same order, so it trails its predecessor BLOCK-WISE — every
// scan publishes how many blocks it has completed, a dependent scan reads that before it loads its next group.  All
// scans of a batch are ONE launch; a warp takes its scan from a ticket counter, and the host lists predecessors
// before their dependents, so a warp only ever waits for warps that started before it (no deadlock, whatever the
// number of resident warps).  The luma chain of libjpeg's standard script (first passes | refinement | last
// refinement) thereby costs its longest scan instead of the sum of the three.
#pragma once
#include "irp_jpeg.cuh"

namespace irp {

constexpr int kProgRing = 2048;     // ring of stream words (8 KB); refilled in halves
constexpr int kProgGroup = 8;       // blocks a refinement scan loads at a time
constexpr int kProgMaxPred = 6;     // predecessors a scan can name (the host falls back to "after everything before it" past that)
constexpr int kProgDone = 0x7FFFFFFF;
// an MCU of a first pass consumes at most 10 blocks x 64 symbols x 31 bits < 640 words, a refinement group at most
// 8 x 63 x 18 bits < 300 words: half a ring is always enough until the next refill point
constexpr int kProgAhead = kProgRing / 2 - 64;   // after a refill: kProgAhead <= staged - next word < kProgRing - 64
static_assert(kProgAhead >= 640, "ring too small for one MCU");

struct ProgScan {                   // one scan of one image
  int img, ns;                      // image, components in the scan
  int comp[3];
  int dc_tab[3], ac_tab[3];         // index into the batch's scan tables, -1: not needed by this scan
  int ss, se, ah, al;
  int restart;                      // MCUs per restart interval at this scan (0: one stream)
  int stream_first, nstreams;
  int npred;                        // earlier scans of the same image whose coefficients this scan reads or overwrites
  int pred[kProgMaxPred];           // their indices in the batch's scan list (always lower than this scan's)
  int pred_blockwise[kProgMaxPred]; // 1: same block order (trail it block by block); 0: wait for its end
};
struct ProgStream {                 // one restart interval of a scan (or the whole scan), un-stuffed, 4-byte aligned
  unsigned long long word_off;      // into the batch data buffer, in 32-bit words
  unsigned int nwords, pad;
};

struct ProgReader {
  uint32_t* ring;
  const uint32_t* src;
  uint32_t nwords;        // words of the stream (zeros are fed past its end, as libjpeg does)
  uint32_t staged;        // words staged so far (warp-uniform)
  unsigned long long acc; // lane 0 only from here
  int avail;
  uint32_t nw;            // next word to merge into acc
  uint32_t ahead;         // ring[nw], loaded when the previous word was merged: the shared-memory latency is off the symbol chain
  __device__ __forceinline__ void open(uint32_t* r, const uint32_t* s, uint32_t n) {
    ring = r; src = s; nwords = n; staged = 0; acc = 0; avail = 0; nw = 0; ahead = 0;
  }
  // warp-converged: keep at least kProgAhead unread words staged.  The ring always holds the 64 words BEFORE the
  // next word to merge as well: up to two of them are still partly unread in lane 0's bit buffer, and the lanes
  // fetch correction bits from there (bit_at) after lane 0 has moved past them.
  __device__ __forceinline__ void refill(int lane) {
    const uint32_t rd = __shfl_sync(0xffffffffu, nw, 0);
    bool any = false;
    while (staged - rd < (uint32_t)kProgAhead) {
      for (int i = lane; i < kProgRing / 2; i += 32) {
        const uint32_t w = staged + i;
        ring[w & (kProgRing - 1)] = w < nwords ? bswap32(__ldg(src + w)) : 0u;
      }
      staged += kProgRing / 2;
      any = true;
    }
    if (any) {
      __syncwarp();
      if (rd == 0 && lane == 0) ahead = ring[0];   // first refill of a stream
    }
  }
  __device__ __forceinline__ void fill() {
    if (avail < 32) {
      acc |= (unsigned long long)ahead << (32 - avail);
      avail += 32;
      nw++;
      ahead = ring[nw & (kProgRing - 1)];
    }
  }
  __device__ __forceinline__ uint32_t top32() const { return (uint32_t)(acc >> 32); }
  __device__ __forceinline__ void skip(int n) {
    acc <<= n;
    avail -= n;
  }
  __device__ __forceinline__ uint32_t bits(int n) {   // 1 <= n <= 32
    fill();
    const uint32_t v = top32() >> (32 - n);
    skip(n);
    return v;
  }
  __device__ __forceinline__ uint32_t bit() {
    fill();
    const uint32_t v = (uint32_t)(acc >> 63);
    skip(1);
    return v;
  }
  __device__ __forceinline__ void skipn(int n) {      // any n >= 0
    if (n <= avail - 32) {                              // the common case: a handful of correction bits
      skip(n);
      return;
    }
    while (n > 0) {
      fill();
      const int t = min(n, 32);
      skip(t);
      n -= t;
    }
  }
  __device__ __forceinline__ uint32_t pos() const { return nw * 32u - (uint32_t)avail; }   // bits consumed so far
  // any lane: bit `b` of the stream, from the ring (valid for positions staged and not yet overwritten)
  __device__ __forceinline__ uint32_t bit_at(uint32_t b) const { return (ring[(b >> 5) & (kProgRing - 1)] >> (31u - (b & 31u))) & 1u; }
};

struct ProgGeom {
  int single;            // one component: the scan walks that component's own blocks
  int mcus_x, mcus;      // MCUs per row / in the scan
  int bpm;               // blocks per MCU of this scan
};

// the compact DC array is in MCU order of the FRAME (irp_jpeg.cuh): block (gy, gx) of component c
__device__ __forceinline__ size_t prog_dc_index(const JpegImg& im, int c, int gy, int gx) {
  const int ch = im.comp_h[c], cv = im.comp_v[c];
  const int my = gy / cv, mx = gx / ch;
  return (size_t)im.dc_off[c] + ((size_t)my * im.mcux + mx) * (ch * cv) + ((gy - my * cv) * ch + (gx - mx * ch));
}

__global__ void __launch_bounds__(32)
prog_scan_kernel(const JpegImg* __restrict__ imgs, const ProgScan* __restrict__ scans, const ProgStream* __restrict__ streams,
                 const HuffDev* __restrict__ tabs, const uint32_t* __restrict__ data, int16_t* __restrict__ coef_arena,
                 int16_t* __restrict__ dcv, int* __restrict__ progress /* [0]: ticket counter, [1 + scan]: blocks completed */) {
  __shared__ uint32_t ring[kProgRing];
  __shared__ HuffDev ht[6];   // DC tables of the scan's components, then their AC tables
  // refinement scans, per block of a group: the history mask and the list of zero positions (built by the warp), and
  // what lane 0 leaves behind — the new coefficients, and the correction-bit SEGMENTS: only where in the band a run of
  // correction bits starts (s_mark) and at which stream bit (s_segpos); the lane that owns a coefficient finds its own
  // bit from those
  __shared__ unsigned long long s_nz[kProgGroup];
  __shared__ uint32_t s_segpos[kProgGroup][64];
  __shared__ __align__(16) uint8_t s_zl[kProgGroup][64];     // band positions of the block's still-zero coefficients, ascending
  __shared__ __align__(16) uint8_t s_mark[kProgGroup][64];   // at the band position where a run of correction bits starts: its index + 1
  __shared__ __align__(16) uint8_t s_new[kProgGroup][64];    // coefficient that becomes nonzero in this scan: 1 positive, 2 negative
  __shared__ int s_nzr[kProgGroup];
  const int lane = threadIdx.x;
  // scans are taken in ticket order: whoever this scan waits for holds a lower ticket, i.e. is running or done
  int ticket = 0;
  if (lane == 0) ticket = atomicAdd(progress, 1);
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  const ProgScan sc = scans[ticket];
  const JpegImg& im = imgs[sc.img];
  volatile int* my_progress = progress + 1 + ticket;
  // wait until every predecessor has completed `need` blocks (or has ended, for one with another block order)
  auto wait_for = [&](int need) {
    if (sc.npred == 0) return;
    if (lane == 0) {
      for (int k = 0; k < sc.npred; k++) {
        const volatile int* pp = progress + 1 + sc.pred[k];
        const int target = sc.pred_blockwise[k] ? need : kProgDone;
        while (*pp < target) __nanosleep(200);
      }
      __threadfence();
    }
    __syncwarp();
  };
  auto publish = [&](int done) {   // all lanes: their stores first
    __threadfence();
    __syncwarp();
    if (lane == 0) *my_progress = done;
  };
  for (int t = 0; t < 6; t++) {
    const int src = t < 3 ? sc.dc_tab[t] : sc.ac_tab[t - 3];
    if ((t % 3) < sc.ns && src >= 0) {
      const uint32_t* s = reinterpret_cast<const uint32_t*>(tabs + src);
      uint32_t* d = reinterpret_cast<uint32_t*>(&ht[t]);
      for (int i = lane; i < (int)(sizeof(HuffDev) / 4); i += 32) d[i] = __ldg(s + i);
    }
  }
  __syncwarp();
  ProgGeom g;
  g.single = sc.ns == 1;
  const int c0 = sc.comp[0];
  if (g.single) {
    g.mcus_x = (im.comp_dw[c0] + 7) >> 3;
    g.mcus = g.mcus_x * ((im.comp_dh[c0] + 7) >> 3);
    g.bpm = 1;
  } else {
    g.mcus_x = im.mcux;
    g.mcus = im.mcux * im.mcuy;
    g.bpm = 0;
    for (int i = 0; i < sc.ns; i++) g.bpm += im.comp_h[sc.comp[i]] * im.comp_v[sc.comp[i]];
  }
  const int per = sc.restart > 0 ? sc.restart : g.mcus;
  ProgReader rd;

  for (int s = 0; s < sc.nstreams; s++) {
    const int m0 = s * per, m1 = min(g.mcus, m0 + per);
    if (m0 >= m1) break;
    const ProgStream st = streams[sc.stream_first + s];
    const uint32_t* src = data + st.word_off;

    // ---------------- DC refinement: block b of the interval owns bit b ----------------
    if (sc.ss == 0 && sc.ah > 0) {
      wait_for(m1);
      const int nb = (m1 - m0) * g.bpm;
      for (int b = lane; b < nb; b += 32) {
        const uint32_t w = (uint32_t)(b >> 5) < st.nwords ? bswap32(__ldg(src + (b >> 5))) : 0u;
        if (!((w >> (31 - (b & 31))) & 1u)) continue;
        const int m = m0 + b / g.bpm;
        int k = b - (b / g.bpm) * g.bpm;
        const int my = m / g.mcus_x, mx = m - my * g.mcus_x;
        for (int i = 0; i < sc.ns; i++) {
          const int c = sc.comp[i];
          const int nh = g.single ? 1 : im.comp_h[c], nv = g.single ? 1 : im.comp_v[c];
          if (k < nh * nv) {
            int16_t* dp = dcv + prog_dc_index(im, c, my * nv + k / nh, mx * nh + k % nh);
            *dp = (int16_t)(__ldcg(dp) | (1 << sc.al));   // written by another warp moments ago: read past L1
            break;
          }
          k -= nh * nv;
        }
      }
      publish(m1);
      continue;
    }

    rd.open(ring, src, st.nwords);
    __syncwarp();
    int eobrun = 0;

    // ---------------- first passes (DC, AC, and the scans of a sequential file): lane 0 decodes and stores ----------------
    if (sc.ah == 0) {
      int pred[3] = {0, 0, 0};
      int my = m0 / g.mcus_x, mx = m0 - my * g.mcus_x;
      for (int m = m0; m < m1; m++) {
        rd.refill(lane);
        if ((m & 15) == 0 || m == m0) {
          wait_for(min((m | 15) + 1, m1));
          if (m > m0) publish(m);   // lane 0's stores of the last sixteen MCUs
        }
        if (lane == 0) {
#pragma unroll 1
          for (int i = 0; i < sc.ns; i++) {
            const int c = sc.comp[i];
            const int nh = g.single ? 1 : im.comp_h[c], nv = g.single ? 1 : im.comp_v[c];
            const HuffDev& dct = ht[i];
            const HuffDev& act = ht[3 + i];
            for (int by = 0; by < nv; by++)
              for (int bx = 0; bx < nh; bx++) {
                const int gy = my * nv + by, gx = mx * nh + bx;
                if (sc.ss == 0) {
                  rd.fill();
                  const uint32_t b32 = rd.top32();
                  int len;
                  const int sz = huff_decode(dct, b32 >> 16, len) & 15;
                  if (sz) pred[i] += huff_extend((int)((b32 << len) >> (32 - sz)), sz);
                  rd.skip(len + sz);
                  dcv[prog_dc_index(im, c, gy, gx)] = (int16_t)(pred[i] * (1 << sc.al));
                }
                if (sc.se > 0) {
                  if (eobrun > 0) {
                    eobrun--;
                    continue;
                  }
                  int16_t* blk = coef_arena + im.coef_off[c] + ((size_t)gy * im.comp_bw[c] + gx) * 64;
                  for (int k = max(sc.ss, 1); k <= sc.se; k++) {
                    rd.fill();
                    const uint32_t b32 = rd.top32();
                    int len;
                    const int rs = huff_decode(act, b32 >> 16, len);
                    const int r = rs >> 4, sz = rs & 15;
                    if (sz) {
                      k += r;
                      const int v = huff_extend((int)((b32 << len) >> (32 - sz)), sz);
                      rd.skip(len + sz);
                      if (k > 63) break;                       // corrupt data only
                      blk[k] = (int16_t)(v * (1 << sc.al));    // the arena keeps zigzag order
                    } else {
                      rd.skip(len);
                      if (r == 15) {
                        k += 15;
                      } else {                                 // end of band for 2^r + extra blocks, this one included
                        eobrun = 1 << r;
                        if (r) eobrun += (int)rd.bits(r);
                        eobrun--;
                        break;
                      }
                    }
                  }
                }
              }
          }
        }
        if (++mx == g.mcus_x) {
          mx = 0;
          my++;
        }
      }
      publish(m1);
      continue;
    }

    // ---------------- AC refinement (one component): groups of eight blocks ----------------
    {
      const int c = c0;
      const int ss = sc.ss, se = sc.se;
      const int p1 = 1 << sc.al;
      const unsigned long long band = (se >= 63 ? ~0ull : ((1ull << (se + 1)) - 1ull)) & ~((1ull << ss) - 1ull);
      int16_t* base = coef_arena + im.coef_off[c];
      const HuffDev& act = ht[3];
      for (int n0 = m0; n0 < m1; n0 += kProgGroup) {
        rd.refill(lane);
        wait_for(min(n0 + kProgGroup, m1));
        int16_t* ptr[kProgGroup];
        int lo[kProgGroup], hi[kProgGroup];
#pragma unroll
        for (int q = 0; q < kProgGroup; q++) {
          const int n = min(n0 + q, m1 - 1);
          const int by = n / g.mcus_x, bx = n - by * g.mcus_x;
          ptr[q] = base + ((size_t)by * im.comp_bw[c] + bx) * 64;
          lo[q] = __ldcg(ptr[q] + lane);        // the predecessor scan wrote these moments ago, from another SM: read past L1
          hi[q] = __ldcg(ptr[q] + lane + 32);
        }
        for (int i = lane; i < kProgGroup * 16; i += 32) {       // last group's marks and new coefficients
          reinterpret_cast<uint32_t*>(&s_mark[0][0])[i] = 0u;
          reinterpret_cast<uint32_t*>(&s_new[0][0])[i] = 0u;
        }
#pragma unroll
        for (int q = 0; q < kProgGroup; q++) {
          const uint32_t a = __ballot_sync(0xffffffffu, lo[q] != 0), b = __ballot_sync(0xffffffffu, hi[q] != 0);
          const unsigned long long nz = (((unsigned long long)b << 32) | a) & band;
          // zero positions of the band in order: a symbol's run of r zeros is then one table look-up for lane 0
          const unsigned long long Z = ~nz & band;
          if ((Z >> lane) & 1ull) s_zl[q][__popcll(Z & ((1ull << lane) - 1ull))] = (uint8_t)lane;
          if ((Z >> (lane + 32)) & 1ull) s_zl[q][__popcll(Z & ((1ull << (lane + 32)) - 1ull))] = (uint8_t)(lane + 32);
          if (lane == 0) {
            s_nz[q] = nz;
            s_nzr[q] = __popcll(Z);
          }
        }
        __syncwarp();
        if (lane == 0) {
          const int cnt = min(kProgGroup, m1 - n0);
#pragma unroll 1
          for (int q = 0; q < cnt; q++) {
            const int nzr = s_nzr[q];
            int nseg = 0, j = 0;   // j: zeros of the band below position k
            int k = ss;
            if (eobrun == 0) {
              while (k <= se) {
                rd.fill();
                const uint32_t b32 = rd.top32();
                int len;
                const int rs = huff_decode(act, b32 >> 16, len);
                const int r = rs >> 4, sz = rs & 15;
                bool neg = false;
                if (sz) {
                  neg = !((b32 << len) >> 31);                 // a newly nonzero coefficient is +-1 at this bit position
                  rd.skip(len + 1);
                } else {
                  rd.skip(len);
                  if (r != 15) {
                    eobrun = 1 << r;
                    if (r) eobrun += (int)rd.bits(r);
                    break;                                     // the rest of the band only takes correction bits
                  }
                }
                // step over r still-zero coefficients: the target is the (r + 1)-th zero from k on; every nonzero
                // coefficient passed on the way takes a correction bit
                int target, zeros;
                if (j + r < nzr) {
                  target = s_zl[q][j + r];
                  zeros = r;
                } else {                                       // band exhausted (corrupt data only)
                  target = se + 1;
                  zeros = nzr - j;
                }
                const int ncorr = target - k - zeros;
                if (ncorr > 0) {
                  s_segpos[q][nseg] = rd.pos();
                  s_mark[q][k] = (uint8_t)(++nseg);
                  rd.skipn(ncorr);
                }
                if (sz && target <= se) s_new[q][target] = neg ? 2 : 1;
                k = target + 1;
                j += zeros + 1;
              }
            }
            if (eobrun > 0) {
              const int ncorr = k <= se ? se + 1 - k - (nzr - min(j, nzr)) : 0;
              if (ncorr > 0) {
                s_segpos[q][nseg] = rd.pos();
                s_mark[q][k] = (uint8_t)(++nseg);
                rd.skipn(ncorr);
              }
              eobrun--;
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < kProgGroup; q++) {
          if (n0 + q >= m1) break;
          const unsigned long long nz = s_nz[q];
          const unsigned long long seg = ((unsigned long long)__ballot_sync(0xffffffffu, s_mark[q][lane + 32] != 0) << 32) |
                                         __ballot_sync(0xffffffffu, s_mark[q][lane] != 0);
#pragma unroll
          for (int half = 0; half < 2; half++) {
            const int pos = lane + 32 * half;
            int v = half ? hi[q] : lo[q];
            const unsigned long long upto = (2ull << pos) - 1ull;        // positions <= pos
            const unsigned long long started = seg & upto;               // segments that begin at or before this coefficient
            bool cb = false;
            if (((nz >> pos) & 1ull) && started) {
              // the coefficient's segment is the last one started; its bit is the segment's first bit plus the number
              // of nonzero coefficients between the segment's start and this one
              const int ks = 63 - __clzll((long long)started);
              const uint32_t rank = (uint32_t)__popcll(nz & (~0ull << ks) & (upto >> 1));
              cb = rd.bit_at(s_segpos[q][s_mark[q][ks] - 1] + rank) != 0u;
            }
            const int nv = s_new[q][pos];
            if (cb && !(v & p1)) v += v >= 0 ? p1 : -p1;
            if (nv) v = nv == 2 ? -p1 : p1;
            if (cb || nv) ptr[q][pos] = (int16_t)v;
          }
        }
        publish(min(n0 + kProgGroup, m1));   // (also the warp barrier between this group's reads of the marks and the next one's clearing)
      }
    }
  }
  publish(kProgDone);
}

}  // namespace irp
