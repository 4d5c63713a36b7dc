// irp_jpeg_enc.cuh — baseline JPEG ENCODE on the device: the `.jpeg({quality: 85, chromaSubsampling: '4:4:4'})`
// stage that ends preprocessImage (server-node/src/middleware/imagePreprocess.js:50-53), so that the resized
// image leaves the GPU as ~1 MB of file bytes instead of 9.4 MB of pixels.
//
// What is reproduced bit for bit is libjpeg-turbo's BASELINE encoder (the library under sharp / libvips and
// under Pillow): jccolor.c RGB -> YCbCr in 16-bit fixed point, jfdctint.c (accurate integer forward DCT),
// jcdctmgr.c quantisation by 16-bit reciprocals, jchuff.c sequential Huffman coding with the Annex K tables or
// with tables optimised per image (optimize_coding), 4:4:4, no restart markers.  mozjpeg's trellis quantisation and progressive scan optimisation (sharp's
// `mozjpeg: true`) are file-size optimisations on top of the same transform; they are NOT reproduced — a
// decoder sees the same kind of image, the file is ~10 % larger.
//
// Stages (all device; the host only sizes buffers from two small read-backs and writes the marker segments):
//   jenc_dct_kernel    one thread per MCU: colour convert, 8x8 forward DCT, quantise; coefficients stored in
//                      zigzag order (int16) + per block {AC code bits, DC value}
//   jenc_blockbits_kernel / seg_scan_kernel   bits per block (DC difference needs the previous block) ->
//                      exclusive scan per image = bit offset of every block, total bits per image
//   jenc_pack_kernel   one thread per block: Huffman-code the block into the image's bit stream at its offset
//                      (whole words stored, the words shared with a neighbour block OR-ed atomically)
//   jenc_hist_kernel / jenc_blocklen_kernel   (IRP_JPEG_OPTIMIZE) symbol statistics per image for the host's
//                      jpeg_gen_optimal_table, then the block lengths under the image's own tables
//   jenc_ffcount_kernel / seg_scan_kernel / jenc_stuff_kernel   0xFF -> 0xFF00 byte stuffing as a stream expansion
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace irp {

constexpr int kEncThreads = 128;
constexpr int kEncPitchH = 66;            // halfwords per thread in the zigzag scratch (33 words: conflict-free)
constexpr int kStuffChunk = 64;           // bytes of raw bit stream per stuffing thread
constexpr int kScanThreads = 1024;

struct EncTables {                        // one quality setting
  uint32_t q[2][64];                      // natural order: reciprocal (16) | correction << 16 (11) | shift << 27
  uint32_t dc[2][16];                     // Huffman code | length << 16, by category
  uint32_t ac[2][256];                    // by (run << 4) | size
};

struct EncImg {                           // one image of an encode launch
  const uint8_t* px;
  unsigned long long pitch;
  int w, h, c, bw, bh, nblk;              // nblk = bw * bh * c blocks in scan order (MCU-interleaved)
  unsigned long long blk0;                // first block in the launch-wide block arrays
  unsigned long long raw_off;             // byte offset of this image's un-stuffed bit stream (4-aligned)
  unsigned long long raw_bytes;           // its length in bytes (ceil(total bits / 8)), known after the first scan
  unsigned long long chunk0;              // first stuffing chunk in the launch-wide chunk arrays
  unsigned long long out_off;             // byte offset of the stuffed stream
  int nchunks, reserved;
};

__device__ __forceinline__ uint32_t enc_bswap(uint32_t x) { return __byte_perm(x, 0u, 0x0123); }

__constant__ uint8_t c_enc_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                         41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                         30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// jfdctint.c jpeg_fdct_islow, one 1-D pass over 8 values at stride S.  FIRST: the row pass (results scaled up by
// PASS1_BITS); otherwise the column pass (PASS1_BITS removed again).
__device__ __forceinline__ int enc_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
template <bool FIRST, int S>
__device__ __forceinline__ void fdct_1d(int* d) {
  const int t0 = d[0] + d[7 * S], t7 = d[0] - d[7 * S], t1 = d[S] + d[6 * S], t6 = d[S] - d[6 * S];
  const int t2 = d[2 * S] + d[5 * S], t5 = d[2 * S] - d[5 * S], t3 = d[3 * S] + d[4 * S], t4 = d[3 * S] - d[4 * S];
  const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
  constexpr int kEven = FIRST ? 13 - 2 : 13 + 2;
  if (FIRST) {
    d[0] = (t10 + t11) << 2;
    d[4 * S] = (t10 - t11) << 2;
  } else {
    d[0] = enc_descale(t10 + t11, 2);
    d[4 * S] = enc_descale(t10 - t11, 2);
  }
  int z1 = (t12 + t13) * 4433;
  d[2 * S] = enc_descale(z1 + t13 * 6270, kEven);
  d[6 * S] = enc_descale(z1 + t12 * (-15137), kEven);
  z1 = t4 + t7;
  int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
  const int z5 = (z3 + z4) * 9633;
  const int a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
  z1 *= -7373;
  z2 *= -20995;
  z3 = z3 * (-16069) + z5;
  z4 = z4 * (-3196) + z5;
  d[7 * S] = enc_descale(a4 + z1 + z3, kEven);
  d[5 * S] = enc_descale(a5 + z2 + z4, kEven);
  d[3 * S] = enc_descale(a6 + z2 + z3, kEven);
  d[S] = enc_descale(a7 + z1 + z4, kEven);
}

__device__ __forceinline__ int enc_nbits(int a) { return 32 - __clz(a); }   // a >= 0

// One thread per 8x8 block of one component (4:4:4: an MCU is one block of each).  grid = (ceil(max blocks / 128), n images).
// The three threads of an MCU read the same pixels (one L1 line) and keep only their own component: 64 live values
// per thread instead of the packed MCU; capped at 80 registers (40 bytes of spill) so that six CTAs share an SM:
// 1.73 ms at 128 registers, 1.48 at 96, 1.32 at 80, 1.37 at 64.
__global__ void __launch_bounds__(kEncThreads, 6)
jenc_dct_kernel(const EncImg* __restrict__ imgs, const EncTables* __restrict__ tabs, int16_t* __restrict__ coef, uint32_t* __restrict__ blkinfo) {
  __shared__ EncTables T;
  __shared__ int16_t zs[kEncThreads * kEncPitchH];
  for (int i = threadIdx.x; i < (int)(sizeof(EncTables) / 4); i += kEncThreads) ((uint32_t*)&T)[i] = ((const uint32_t*)tabs)[i];
  __syncthreads();
  const EncImg& im = imgs[blockIdx.y];
  const int w = im.w, h = im.h, c = im.c, bw = im.bw;
  const int t = blockIdx.x * kEncThreads + threadIdx.x;
  if (t >= im.nblk) return;
  const int m = c == 3 ? t / 3 : t, comp = t - m * c;
  const int bx = m % bw, by = m / bw, x0 = bx * 8, y0 = by * 8;
  const uint8_t* __restrict__ px = im.px;
  const size_t pitch = (size_t)im.pitch;
  int d[64];
  if (c == 3) {
    // jccolor.c rgb_ycc_convert: 16-bit fixed point, the chroma rows biased by 128.5 - 1 LSB; level shift folded in
    const int cr = comp == 0 ? 19595 : (comp == 1 ? -11059 : 32768);
    const int cg = comp == 0 ? 38470 : (comp == 1 ? -21709 : -27439);
    const int cb = comp == 0 ? 7471 : (comp == 1 ? 32768 : -5329);
    const int c0 = comp == 0 ? 32768 : (128 << 16) + 32767;
    const bool fast = x0 + 8 <= w && (((uintptr_t)px | pitch) & 3) == 0;   // whole block inside, rows 4-byte aligned
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const uint8_t* row = px + (size_t)min(y0 + r, h - 1) * pitch;   // edge blocks replicate the last row / column (jcprepct.c)
      if (fast) {
        const uint2* p = (const uint2*)(row + x0 * 3);
        const uint2 a = __ldg(p), b = __ldg(p + 1), cc = __ldg(p + 2);
        const uint32_t raw[6] = {a.x, a.y, b.x, b.y, cc.x, cc.y};
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int R = (raw[(3 * j) >> 2] >> (8 * ((3 * j) & 3))) & 255;
          const int G = (raw[(3 * j + 1) >> 2] >> (8 * ((3 * j + 1) & 3))) & 255;
          const int B = (raw[(3 * j + 2) >> 2] >> (8 * ((3 * j + 2) & 3))) & 255;
          d[r * 8 + j] = ((cr * R + cg * G + cb * B + c0) >> 16) - 128;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const uint8_t* p = row + (size_t)min(x0 + j, w - 1) * 3;
          d[r * 8 + j] = ((cr * (int)__ldg(p) + cg * (int)__ldg(p + 1) + cb * (int)__ldg(p + 2) + c0) >> 16) - 128;
        }
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const uint8_t* row = px + (size_t)min(y0 + r, h - 1) * pitch;
#pragma unroll
      for (int j = 0; j < 8; j++) d[r * 8 + j] = (int)__ldg(row + min(x0 + j, w - 1)) - 128;
    }
  }
#pragma unroll
  for (int r = 0; r < 8; r++) fdct_1d<true, 1>(d + r * 8);
#pragma unroll
  for (int j = 0; j < 8; j++) fdct_1d<false, 8>(d + j);
  // jcdctmgr.c quantize(): |x| + correction, times the 16-bit reciprocal, shifted; sign restored
  int16_t* mine = zs + threadIdx.x * kEncPitchH;
  const uint32_t* qt = T.q[comp ? 1 : 0];
#pragma unroll
  for (int i = 0; i < 64; i++) {
    const uint32_t e = qt[i];
    const int v = d[i];
    const uint32_t prod = ((uint32_t)abs(v) + ((e >> 16) & 0x7FFu)) * (e & 0xFFFFu);
    const int qv = (int)(prod >> ((e >> 27) + 16));
    mine[i] = (int16_t)(v < 0 ? -qv : qv);
  }
  // zigzag read-back: 8 coefficients per 16-byte store; AC code lengths summed on the way (jchuff.c encode_one_block)
  const uint32_t* act = T.ac[comp ? 1 : 0];
  const size_t g = (size_t)im.blk0 + t;
  uint4* dst = (uint4*)(coef + g * 64);
  int run = 0, bits = 0, dcv = 0;
#pragma unroll 1
  for (int k8 = 0; k8 < 8; k8++) {
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int v = mine[c_enc_zigzag[k8 * 8 + i]];
      if (i & 1) pk[i >> 1] |= (uint32_t)(uint16_t)v << 16;
      else pk[i >> 1] = (uint16_t)v;
      if (k8 == 0 && i == 0) {
        dcv = v;
      } else if (v == 0) {
        run++;
      } else {
        bits += (run >> 4) * (int)(act[0xF0] >> 16);
        const int nb = enc_nbits(abs(v));
        bits += (int)(act[((run & 15) << 4) | nb] >> 16) + nb;
        run = 0;
      }
    }
    dst[k8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  if (run) bits += (int)(act[0] >> 16);
  blkinfo[g] = (uint32_t)bits | ((uint32_t)(uint16_t)dcv << 16);
}

// bits of every block = AC bits + the DC difference code (previous block of the same component, 0 at the start)
__global__ void jenc_blockbits_kernel(const EncImg* __restrict__ imgs, const EncTables* __restrict__ tabs, const uint32_t* __restrict__ blkinfo,
                                      uint32_t* __restrict__ bits) {
  const EncImg& im = imgs[blockIdx.y];
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= im.nblk) return;
  const size_t a = (size_t)im.blk0 + g;
  const uint32_t e = __ldg(blkinfo + a);
  const int prev = g >= im.c ? (int)(int16_t)(__ldg(blkinfo + a - im.c) >> 16) : 0;
  const int diff = (int)(int16_t)(e >> 16) - prev;
  const int nb = enc_nbits(abs(diff));
  bits[a] = (e & 0xFFFFu) + (__ldg(&tabs->dc[(g % im.c) ? 1 : 0][nb]) >> 16) + nb;
}

// In-place exclusive scan of one u32 segment per CTA (segment s = [start[s], start[s] + len[s])), total to totals[s].
struct ScanSeg {
  unsigned long long start;
  unsigned long long len;
};
__global__ void __launch_bounds__(kScanThreads)
seg_scan_kernel(const ScanSeg* __restrict__ segs, uint32_t* __restrict__ vals, unsigned long long* __restrict__ totals) {
  __shared__ unsigned long long warp_tot[32];
  __shared__ unsigned long long carry_s;
  const ScanSeg sg = segs[blockIdx.x];
  uint32_t* v = vals + sg.start;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long carry = 0;
  for (unsigned long long base = 0; base < sg.len; base += kScanThreads * 4) {
    const unsigned long long i0 = base + (unsigned long long)threadIdx.x * 4;
    uint32_t x[4];
#pragma unroll
    for (int k = 0; k < 4; k++) x[k] = i0 + k < sg.len ? v[i0 + k] : 0u;
    const unsigned long long mine = (unsigned long long)x[0] + x[1] + x[2] + x[3];
    unsigned long long inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      unsigned long long t = warp_tot[lane], ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += y;
      }
      warp_tot[lane] = ti - t;
      if (lane == 31) carry_s = ti;
    }
    __syncthreads();
    unsigned long long acc = carry + warp_tot[wid] + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (i0 + k < sg.len) v[i0 + k] = (uint32_t)acc;
      acc += x[k];
    }
    carry += carry_s;
    __syncthreads();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

// ---- optimised Huffman tables (libjpeg's optimize_coding): symbol statistics per image, tables on the host ----
struct EncHuff {                          // per image: codes (| length << 16) in the pack pass, counts in the histogram pass
  uint32_t dc[2][16];
  uint32_t ac[2][256];
};

// jchuff.c htest_one_block for every block: one thread per block, counts gathered per CTA in shared memory
__global__ void __launch_bounds__(kEncThreads)
jenc_hist_kernel(const EncImg* __restrict__ imgs, const int16_t* __restrict__ coef, const uint32_t* __restrict__ blkinfo, EncHuff* __restrict__ hist) {
  __shared__ EncHuff h;
  for (int i = threadIdx.x; i < (int)(sizeof(EncHuff) / 4); i += kEncThreads) ((uint32_t*)&h)[i] = 0u;
  __syncthreads();
  const EncImg& im = imgs[blockIdx.y];
  const int g = blockIdx.x * kEncThreads + threadIdx.x;
  if (g < im.nblk) {
    const size_t a = (size_t)im.blk0 + g;
    const int t = (g % im.c) ? 1 : 0;
    const int prev = g >= im.c ? (int)(int16_t)(__ldg(blkinfo + a - im.c) >> 16) : 0;
    const uint4* src = (const uint4*)(coef + a * 64);
    int run = 0;
#pragma unroll 1
    for (int k8 = 0; k8 < 8; k8++) {
      const uint4 q = __ldg(src + k8);
      const uint32_t pk[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const int v = (int)(int16_t)(pk[i >> 1] >> ((i & 1) * 16));
        if (k8 == 0 && i == 0) {
          atomicAdd(&h.dc[t][enc_nbits(abs(v - prev))], 1u);
        } else if (v == 0) {
          run++;
        } else {
          if (run > 15) atomicAdd(&h.ac[t][0xF0], (uint32_t)(run >> 4));
          atomicAdd(&h.ac[t][((run & 15) << 4) | enc_nbits(abs(v))], 1u);
          run = 0;
        }
      }
    }
    if (run) atomicAdd(&h.ac[t][0], 1u);
  }
  __syncthreads();
  uint32_t* out = (uint32_t*)(hist + blockIdx.y);
  for (int i = threadIdx.x; i < (int)(sizeof(EncHuff) / 4); i += kEncThreads) {
    const uint32_t v = ((uint32_t*)&h)[i];
    if (v) atomicAdd(out + i, v);
  }
}

// bits of every block under the image's own tables (replaces the Annex K lengths jenc_dct_kernel summed)
__global__ void __launch_bounds__(kEncThreads)
jenc_blocklen_kernel(const EncImg* __restrict__ imgs, const EncHuff* __restrict__ own, const int16_t* __restrict__ coef,
                     const uint32_t* __restrict__ blkinfo, uint32_t* __restrict__ bits) {
  __shared__ EncHuff h;
  for (int i = threadIdx.x; i < (int)(sizeof(EncHuff) / 4); i += kEncThreads) ((uint32_t*)&h)[i] = ((const uint32_t*)(own + blockIdx.y))[i];
  __syncthreads();
  const EncImg& im = imgs[blockIdx.y];
  const int g = blockIdx.x * kEncThreads + threadIdx.x;
  if (g >= im.nblk) return;
  const size_t a = (size_t)im.blk0 + g;
  const int t = (g % im.c) ? 1 : 0;
  const int prev = g >= im.c ? (int)(int16_t)(__ldg(blkinfo + a - im.c) >> 16) : 0;
  const uint4* src = (const uint4*)(coef + a * 64);
  int run = 0, n = 0;
#pragma unroll 1
  for (int k8 = 0; k8 < 8; k8++) {
    const uint4 q = __ldg(src + k8);
    const uint32_t pk[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int v = (int)(int16_t)(pk[i >> 1] >> ((i & 1) * 16));
      if (k8 == 0 && i == 0) {
        const int nb = enc_nbits(abs(v - prev));
        n += (int)(h.dc[t][nb] >> 16) + nb;
      } else if (v == 0) {
        run++;
      } else {
        n += (run >> 4) * (int)(h.ac[t][0xF0] >> 16);
        const int nb = enc_nbits(abs(v));
        n += (int)(h.ac[t][((run & 15) << 4) | nb] >> 16) + nb;
        run = 0;
      }
    }
  }
  if (run) n += (int)(h.ac[t][0] >> 16);
  bits[a] = (uint32_t)n;
}

// bit writer of one block: MSB-first into 32-bit words, byte-swapped on the way out
struct BitOut {
  uint32_t* wp;
  uint32_t cur;
  int fill;
  bool first;
  __device__ __forceinline__ void flush() {
    if (first) atomicOr(wp, enc_bswap(cur));   // shared with the block before
    else *wp = enc_bswap(cur);
    first = false;
    wp++;
  }
  __device__ __forceinline__ void put(uint32_t v, int len) {   // len in 1..27, v < 2^len
    const int room = 32 - fill;
    if (len < room) {
      cur |= v << (room - len);
      fill += len;
    } else {
      const int hi = len - room;
      cur |= v >> hi;
      flush();
      cur = hi ? v << (32 - hi) : 0u;
      fill = hi;
    }
  }
};

// One thread per block.  grid = (ceil(max blocks / 128), n images).  The bit stream area must be zero.
__global__ void __launch_bounds__(kEncThreads)
jenc_pack_kernel(const EncImg* __restrict__ imgs, const EncTables* __restrict__ tabs, const int16_t* __restrict__ coef,
                 const uint32_t* __restrict__ blkinfo, const uint32_t* __restrict__ bitoff, uint8_t* __restrict__ rawbits,
                 const EncHuff* __restrict__ own) {
  __shared__ uint32_t s_dc[2][16], s_ac[2][256];
  const uint32_t* dcs = own ? &own[blockIdx.y].dc[0][0] : &tabs->dc[0][0];   // the image's optimised tables, or Annex K
  const uint32_t* acs = own ? &own[blockIdx.y].ac[0][0] : &tabs->ac[0][0];
  for (int i = threadIdx.x; i < 32; i += kEncThreads) (&s_dc[0][0])[i] = dcs[i];
  for (int i = threadIdx.x; i < 512; i += kEncThreads) (&s_ac[0][0])[i] = acs[i];
  __syncthreads();
  const EncImg& im = imgs[blockIdx.y];
  const int g = blockIdx.x * kEncThreads + threadIdx.x;
  if (g >= im.nblk) return;
  const size_t a = (size_t)im.blk0 + g;
  const int t = (g % im.c) ? 1 : 0;
  const uint32_t off = __ldg(bitoff + a);
  BitOut bo;
  bo.wp = (uint32_t*)(rawbits + im.raw_off) + (off >> 5);
  bo.cur = 0;
  bo.fill = (int)(off & 31);
  bo.first = true;
  const int prev = g >= im.c ? (int)(int16_t)(__ldg(blkinfo + a - im.c) >> 16) : 0;
  const uint4* src = (const uint4*)(coef + a * 64);
  const uint32_t* act = s_ac[t];
  const uint32_t zrl = act[0xF0];
  int run = 0;
#pragma unroll 1
  for (int k8 = 0; k8 < 8; k8++) {
    const uint4 q = __ldg(src + k8);
    const uint32_t pk[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int v = (int)(int16_t)(pk[i >> 1] >> ((i & 1) * 16));
      if (k8 == 0 && i == 0) {
        const int diff = v - prev;
        const int nb = enc_nbits(abs(diff));
        const uint32_t e = s_dc[t][nb];
        const uint32_t extra = (uint32_t)(diff < 0 ? diff - 1 : diff) & ((1u << nb) - 1u);
        bo.put(((e & 0xFFFFu) << nb) | extra, (int)(e >> 16) + nb);
      } else if (v == 0) {
        run++;
      } else {
        while (run > 15) {
          bo.put(zrl & 0xFFFFu, (int)(zrl >> 16));
          run -= 16;
        }
        const int nb = enc_nbits(abs(v));
        const uint32_t e = act[(run << 4) | nb];
        const uint32_t extra = (uint32_t)(v < 0 ? v - 1 : v) & ((1u << nb) - 1u);
        bo.put(((e & 0xFFFFu) << nb) | extra, (int)(e >> 16) + nb);
        run = 0;
      }
    }
  }
  if (run) bo.put(act[0] & 0xFFFFu, (int)(act[0] >> 16));
  if (g == im.nblk - 1) {            // jchuff.c flush_bits: pad the last byte with 1-bits
    const int pad = (8 - (bo.fill & 7)) & 7;
    if (pad) bo.put((1u << pad) - 1u, pad);
  }
  if (bo.fill) atomicOr(bo.wp, enc_bswap(bo.cur));
}

// 0xFF bytes per 64-byte chunk of every raw stream.  grid = (ceil(max chunks / 256), n images)
__global__ void jenc_ffcount_kernel(const EncImg* __restrict__ imgs, const uint8_t* __restrict__ rawbits, uint32_t* __restrict__ cnt) {
  const EncImg& im = imgs[blockIdx.y];
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= im.nchunks) return;
  const uint4* p = (const uint4*)(rawbits + im.raw_off + (size_t)ch * kStuffChunk);
  int n = 0;   // bytes past the end of the stream are zero (the area was cleared), never 0xFF
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint4 q = __ldg(p + k);
    const uint32_t wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint32_t y = wv[j] & (wv[j] >> 1);   // bit 0 of each byte ends up as the AND of its 8 bits
      y &= y >> 2;
      y &= y >> 4;
      n += __popc(y & 0x01010101u);
    }
  }
  cnt[im.chunk0 + ch] = (uint32_t)n;
}

// copy each chunk to its place in the stuffed stream, a zero byte after every 0xFF
__global__ void jenc_stuff_kernel(const EncImg* __restrict__ imgs, const uint8_t* __restrict__ rawbits, const uint32_t* __restrict__ ffoff,
                                  uint8_t* __restrict__ out) {
  const EncImg& im = imgs[blockIdx.y];
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= im.nchunks) return;
  const uint4* p = (const uint4*)(rawbits + im.raw_off + (size_t)ch * kStuffChunk);
  const unsigned long long begin = (unsigned long long)ch * kStuffChunk;
  const int nbytes = (int)min((unsigned long long)kStuffChunk, im.raw_bytes - begin);
  uint8_t* o = out + im.out_off + begin + __ldg(ffoff + im.chunk0 + ch);
#pragma unroll 1
  for (int k = 0; k < 4; k++) {
    const uint4 q = __ldg(p + k);
    const uint32_t wv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 16; j++) {
      if (k * 16 + j < nbytes) {
        const uint32_t b = (wv[j >> 2] >> (8 * (j & 3))) & 255u;
        *o++ = (uint8_t)b;
        if (b == 255u) *o++ = 0;
      }
    }
  }
}

}  // namespace irp
