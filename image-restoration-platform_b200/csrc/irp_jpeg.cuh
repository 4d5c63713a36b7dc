// irp_jpeg.cuh — baseline JPEG decode on the device (sm_100a): the decoder that sits in front of every
// sharp pipeline of the hot path (reference: server-node/src/services/classifier.js:51-52,107,135,199,296
// `sharp(imageBuffer)`; server-node/src/middleware/imagePreprocess.js:40-42), so
// that compressed bytes instead of raw pixels cross PCIe.  Pixel-exact with libjpeg-turbo's default
// decode (JDCT_ISLOW, fancy upsampling), the decoder inside libvips; the parity tests compare it with a CPU
// restatement that is itself held bit-exact against libjpeg-turbo (tests/test_jpeg_*.py).
//
// Stages (one launch each over the whole batch; the host has parsed the headers and removed the byte
// stuffing while it copied the entropy-coded segments into one upload buffer):
//   huffman   Huffman decoding is sequential by nature: a code can only be found once the previous one
//             ended.  It is parallelised the self-synchronising way: the stream is cut into 2048-bit
//             subsequences, every thread decodes its own from the guess "a block starts here", and
//             JPEG's code tables make a wrong guess fall into step with the true decode within a few
//             symbols.  `huff_sync_kernel` iterates state(i) = f_i(state(i-1)) until nothing changes —
//             then every subsequence's start state is the true one — an exclusive scan over the
//             coefficient slots each subsequence advanced gives its place in the output, and
//             `huff_write_kernel` decodes once more, writing coefficients — in ZIGZAG order, 64 int16
//             per block, DC differences apart in a compact MCU-order array.  Restart intervals, when a
//             file has them, are independent streams with known start states.
//   dc        the DC coefficients are coded as differences: one running sum per component and stream,
//             over the compact array.
//   idct      dequantise + jidctint (13-bit constants, two passes), one thread per block in 64 registers
//             (zigzag -> natural order is register renaming), u8 planes.
//   colour    fancy (triangle) chroma upsampling h2v1 / h2v2 / h1v2 + YCbCr -> RGB (16-bit fixed point),
//             written as interleaved u8 rows with a 16-byte pitch: the classify / resize kernels' input.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace irp {

constexpr int kSubBits = 2048;         // bits per subsequence
constexpr int kHuffThreads = 128;      // subsequences per CTA (a CTA never spans two images)
constexpr int kLutBits = 9;

struct HuffDev {                       // one Huffman table
  uint16_t lut[1 << kLutBits];         // (length << 8) | symbol for codes of <= 9 bits, 0 otherwise
  int32_t maxcode[18];                 // canonical decode for longer codes
  int32_t valoff[17];                  // valptr[l] - mincode[l]
  uint8_t vals[256];
};

struct JpegImg {                       // one image of a decode launch (device copy)
  int w, h, ncomp, hmax, vmax, mcux, mcuy, bpm;
  int comp_h[3], comp_v[3], comp_bw[3], comp_bh[3], comp_dw[3], comp_dh[3];
  int blk_comp[8], blk_bx[8], blk_by[8];   // block k of an MCU: component, offset inside the MCU (in blocks)
  int dc_tab[3], ac_tab[3];                // slot (0..5) of the component's DC / AC table among the image's tables
  uint16_t q[3][64];                       // natural order
  int restart;                             // MCUs per restart interval (0: none)
  int nstreams;                            // restart intervals (1 if none)
  int stream_base;                         // first entry of this image in the stream arrays
  int sub_base, nsub;                      // first subsequence (multiple of kHuffThreads), count
  unsigned long long coef_off[3];          // int16 offsets into the coefficient arena
  unsigned long long plane_off[3];         // byte offsets into the plane arena
  unsigned long long dc_off[3];            // first entry of the component in the compact DC array (MCU order)
  unsigned long long out_off;              // byte offset of the RGB / grey output in the pixel arena
  unsigned long long out_pitch;
  int huff_base;                           // first of this image's kHuffSlots tables
  int fancy;                               // fancy (triangle) upsampling; libjpeg drops it when the smallest IDCT is 1x1
  // the decode scale (libjpeg scale 1 / 1, 2, 4, 8 — shrink-on-load): output size, and per component the IDCT size
  // (8 = full), the size of its plane in samples and the expansion the upsampler still has to do.  The entropy-coded
  // data does not know about the scale: the Huffman / scan kernels keep using comp_dw / comp_dh / comp_bw / comp_bh.
  int ow, oh;
  int comp_S[3], comp_sdw[3], comp_sdh[3], comp_hx[3], comp_vx[3];
  int pad;
};

struct JpegStream {                        // one restart interval (or the whole scan)
  unsigned long long bit_off;              // first bit in the unstuffed batch buffer (multiple of 32)
  unsigned long long nbits;
  int img;
  int first_mcu, n_mcu;
  int sub_first;                           // first subsequence (global index), and how many
  int nsub;
  int pad;
};

struct SubState {                          // state AFTER a subsequence: where the next symbol starts
  uint32_t p;                              // bit position relative to the stream
  uint32_t slot;                           // coefficient slot modulo (64 * bpm): block in MCU * 64 + zigzag index
};

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0u, 0x0123); }

// A left-aligned 64-bit bit buffer over a big-endian stream of 32-bit words: seek() once per run, then per symbol
// fill() (one predicated word load when fewer than 32 bits are left), top32() and skip().  One 32-bit look
// covers a Huffman code (<= 16 bits) and its value bits (<= 15), so the dependent chain per symbol is
// "buffer -> LUT -> shift", with no position -> word arithmetic.  The words of a subsequence (and 16 bytes
// beyond: a symbol that starts in it may end there) are staged in shared memory by its owner thread with
// independent 128-bit loads before any decoding, so the chain never waits on global memory.
constexpr int kSubWords = kSubBits / 32 + 4;
struct BitWin {
  const uint32_t* d;  // staged words of ONE subsequence
  uint32_t bit0;      // stream-relative bit position of its first word
  unsigned long long acc;
  int avail, nw;      // valid bits in acc; index of the next word to load
  __device__ __forceinline__ void seek(uint32_t p) {
    const uint32_t q = p - bit0;
    const int wi = min((int)(q >> 5), kSubWords - 2);        // (a position past the staged words only occurs on corrupt data)
    acc = (((unsigned long long)d[wi] << 32) | d[wi + 1]) << (q & 31);
    avail = 64 - (int)(q & 31);
    nw = wi + 2;
  }
  __device__ __forceinline__ void fill() {
    if (avail < 32) {
      acc |= (unsigned long long)d[min(nw, kSubWords - 1)] << (32 - avail);
      avail += 32;
      nw++;
    }
  }
  __device__ __forceinline__ uint32_t top32() const { return (uint32_t)(acc >> 32); }
  __device__ __forceinline__ void skip(int n) {
    acc <<= n;
    avail -= n;
  }
};
// the same buffer straight over global memory (read-only path, L1-cached): the synchronisation launches
// decode little after the first sweep, and staging would cost them their occupancy
struct BitWinG {
  const uint32_t* d;   // the batch buffer
  unsigned long long base;   // stream_bit0
  unsigned long long acc;
  int avail;
  const uint32_t* nw;
  __device__ __forceinline__ void seek(uint32_t p) {
    const unsigned long long q = base + p;
    const uint32_t* w = d + (q >> 5);
    acc = (((unsigned long long)bswap32(__ldg(w)) << 32) | bswap32(__ldg(w + 1))) << (unsigned)(q & 31);
    avail = 64 - (int)(q & 31);
    nw = w + 2;
  }
  __device__ __forceinline__ void fill() {
    if (avail < 32) {
      acc |= (unsigned long long)bswap32(__ldg(nw)) << (32 - avail);
      avail += 32;
      nw++;
    }
  }
  __device__ __forceinline__ uint32_t top32() const { return (uint32_t)(acc >> 32); }
  __device__ __forceinline__ void skip(int n) {
    acc <<= n;
    avail -= n;
  }
};
// stage subsequence `local` of a stream: words are byte-swapped once here
__device__ __forceinline__ void stage_sub(uint32_t* dst, const uint32_t* __restrict__ data, unsigned long long stream_bit0, int local) {
  const uint4* src = reinterpret_cast<const uint4*>(data + (stream_bit0 >> 5) + (size_t)local * (kSubBits / 32));
  // stream starts are 4-byte aligned only: 128-bit loads need 16; fall back to words when they are not
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
#pragma unroll 4
    for (int k = 0; k < kSubWords / 4; k++) {
      const uint4 v = __ldg(src + k);
      dst[4 * k] = bswap32(v.x); dst[4 * k + 1] = bswap32(v.y); dst[4 * k + 2] = bswap32(v.z); dst[4 * k + 3] = bswap32(v.w);
    }
  } else {
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
    for (int k = 0; k < kSubWords; k++) dst[k] = bswap32(__ldg(s32 + k));
  }
}

// three components use at most three DC and three AC tables: six slots instead of the eight a file may define keep
// the write kernel (35 KB of staged stream per CTA) at five CTAs per SM
constexpr int kHuffSlots = 6;
struct HuffSmem {
  HuffDev t[kHuffSlots];
};

__device__ __forceinline__ int huff_decode(const HuffDev& t, uint32_t bits16, int& len) {
  const uint32_t e = t.lut[bits16 >> (16 - kLutBits)];
  if (e) {
    len = (int)(e >> 8);
    return (int)(e & 0xFF);
  }
  int l = kLutBits + 1;
  int code = (int)(bits16 >> (16 - l));
  while (l <= 16 && code > t.maxcode[l]) {
    l++;
    code = (int)(bits16 >> (16 - l));
  }
  if (l > 16) {   // not a code (only reachable while out of sync): consume one bit, harmless symbol
    len = 1;
    return 0;
  }
  len = l;
  return t.vals[(code + t.valoff[l]) & 255];
}

__device__ __forceinline__ int huff_extend(int x, int s) { return x < (1 << (s - 1)) ? x - (1 << s) + 1 : x; }

// Decode the symbols that START in [p, p_end) of one stream, from slot state `slot` (mod 64 * bpm).
// WRITE: store coefficients (in zigzag order, DC differences to the compact array); `abs_blk` is the absolute block index (MCU order) the
// run starts in and `blk_limit` one past the stream's last block.  The block's address is worked out
// once per block, not per coefficient.
template <bool WRITE, class Reader>
__device__ __forceinline__ void huff_run(const JpegImg& im, const HuffSmem& hs, Reader& bw, unsigned long long stream_bit0,
                                         unsigned long long stream_bits, uint32_t& p, uint32_t p_end, uint32_t& slot, uint32_t& advanced,
                                         int16_t* __restrict__ const* coef, int16_t* __restrict__ dcv, uint32_t abs_blk,
                                         uint32_t blk_limit) {
  const uint32_t period = 64u * (uint32_t)im.bpm;
  const uint32_t stop = (uint32_t)min((unsigned long long)p_end, stream_bits);
  if (p >= stop) return;
  bw.seek(p);
  // block by block: the component, its two tables and (when writing) the block's address are looked up once
  // per block; the inner loop over the AC symbols touches one table and the bit buffer only
  while (p < stop && (!WRITE || abs_blk < blk_limit)) {
    uint32_t z = slot & 63u;
    const uint32_t k = slot >> 6;
    const int comp = im.blk_comp[k];
    const HuffDev& dct = hs.t[im.dc_tab[comp]];
    const HuffDev& act = hs.t[im.ac_tab[comp]];
    int16_t* blk_ptr = nullptr;
    int16_t* dc_ptr = nullptr;     // DC differences go to a compact array in MCU order: the prediction scan runs over
    if (WRITE) {                   // contiguous values instead of one 2-byte access per 128-byte block
      const uint32_t mcu = abs_blk / (uint32_t)im.bpm;
      const uint32_t my = mcu / (uint32_t)im.mcux, mx = mcu - my * (uint32_t)im.mcux;
      const int ch = im.comp_h[comp], cv = im.comp_v[comp];
      blk_ptr = coef[comp] + ((size_t)(my * cv + im.blk_by[k]) * im.comp_bw[comp] + (mx * ch + im.blk_bx[k])) * 64;
      dc_ptr = dcv + im.dc_off[comp] + (size_t)mcu * (ch * cv) + (im.blk_by[k] * ch + im.blk_bx[k]);
    }
    const uint32_t z_in = z;
    if (z == 0) {
      bw.fill();
      const uint32_t b32 = bw.top32();
      int len;
      const int s = huff_decode(dct, b32 >> 16, len) & 15;
      if (WRITE) *dc_ptr = (int16_t)(s ? huff_extend((int)((b32 << len) >> (32 - s)), s) : 0);
      bw.skip(len + s);
      p += len + s;
      z = 1;
    }
    while (z < 64u && p < stop) {
      bw.fill();
      const uint32_t b32 = bw.top32();
      int len;
      const int rs = huff_decode(act, b32 >> 16, len);
      const int r = rs >> 4, s = rs & 15;
      if (s == 0) {
        z = (r == 15) ? min(z + 16u, 64u) : 64u;          // ZRL : EOB
      } else {
        uint32_t zz = z + r;
        if (zz > 63u) zz = 63u;                           // only while out of sync / on corrupt data
        if (WRITE) blk_ptr[zz] = (int16_t)huff_extend((int)((b32 << len) >> (32 - s)), s);   // the arena keeps ZIGZAG order
        z = zz + 1;
      }
      bw.skip(len + s);
      p += len + s;
    }
    const uint32_t adv = z - z_in;
    slot += adv;
    if (slot >= period) slot -= period;
    advanced += adv;
    if (WRITE && z >= 64u) abs_blk++;
  }
}

__device__ __forceinline__ void load_huff(HuffSmem& hs, const HuffDev* __restrict__ tabs, int base) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(tabs + base);
  uint32_t* d = reinterpret_cast<uint32_t*>(&hs);
  for (int i = threadIdx.x; i < (int)(sizeof(HuffSmem) / 4); i += blockDim.x) d[i] = __ldg(s + i);
}

// which stream of the image does subsequence `sub` (global index) belong to: streams are in order
__device__ __forceinline__ int find_stream(const JpegStream* __restrict__ streams, int lo, int n, int sub) {
  int a = lo, b = lo + n - 1;
  while (a < b) {
    const int m = (a + b + 1) >> 1;
    if (__ldg(&streams[m].sub_first) <= sub) a = m; else b = m - 1;
  }
  return a;
}

// The fixed-point iteration.  in_state[i] is the start state subsequence i was last decoded from and
// state[i] the state it ended in.  first == 1: every subsequence starts from the guess "a DC symbol of block
// 0 starts at my first bit" (true for the first subsequence of a stream).  Later launches run `rounds` rounds:
// a thread re-decodes only when its predecessor's end state differs from the start state it used, and a CTA
// barrier between rounds lets a correction travel up to `rounds` subsequences per launch.  A launch in which
// nobody re-decoded leaves `changed` at 0: state[i] == f_i(state[i-1]) everywhere, i.e. all states are true.
__global__ void __launch_bounds__(kHuffThreads)
huff_sync_kernel(const JpegImg* __restrict__ imgs, const int* __restrict__ cta_img, const JpegStream* __restrict__ streams,
                 const HuffDev* __restrict__ tabs, const uint32_t* __restrict__ data, SubState* __restrict__ state,
                 SubState* __restrict__ in_state, uint32_t* __restrict__ advanced, int first, int rounds, int* __restrict__ changed) {
  __shared__ HuffSmem hs;
  __shared__ JpegImg im;
  const int img = cta_img[blockIdx.x];
  for (int i = threadIdx.x; i < (int)(sizeof(JpegImg) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(&im)[i] = reinterpret_cast<const uint32_t*>(imgs + img)[i];
  load_huff(hs, tabs, imgs[img].huff_base);
  __syncthreads();
  const int sub0 = blockIdx.x * kHuffThreads;
  const int sub_end = im.sub_base + im.nsub;
  volatile unsigned long long* vstate = reinterpret_cast<volatile unsigned long long*>(state);
  unsigned long long* vin = reinterpret_cast<unsigned long long*>(in_state);
  __shared__ int queue[kHuffThreads];
  __shared__ int qn;

  for (int round = 0; round < rounds; round++) {
    // which subsequences of this CTA must be decoded (again)?  Their indices are packed into a queue so
    // that the decoding threads fill whole warps: a warp with one busy lane costs as much as a full one.
    if (threadIdx.x == 0) qn = 0;
    __syncthreads();
    const int mine = sub0 + threadIdx.x;
    bool need = false;
    if (mine < sub_end) {
      if (first) {
        need = true;
      } else {
        const JpegStream s0 = streams[find_stream(streams, im.stream_base, im.nstreams, mine)];
        need = mine > s0.sub_first && vstate[mine - 1] != vin[mine];
      }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, need);
    int base = 0;
    if ((threadIdx.x & 31) == 0 && ballot) base = atomicAdd(&qn, __popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (need) queue[base + __popc(ballot & ((1u << (threadIdx.x & 31)) - 1u))] = mine;
    __syncthreads();
    const int n = qn;
    if (n == 0) break;     // uniform: nothing to do here now (a change arriving from the previous CTA is caught by the next launch)
    if ((int)threadIdx.x < n) {
      const int sub = queue[threadIdx.x];
      const JpegStream st = streams[find_stream(streams, im.stream_base, im.nstreams, sub)];
      const int local = sub - st.sub_first;
      unsigned long long start;
      if (local == 0)
        start = 0;
      else if (first)
        start = (unsigned long long)((uint32_t)local * kSubBits);
      else
        start = vstate[sub - 1];
      uint32_t p = (uint32_t)start, slot = (uint32_t)(start >> 32), adv = 0;
      BitWinG bw{data, st.bit_off, 0ull, 0, nullptr};
      huff_run<false>(im, hs, bw, st.bit_off, st.nbits, p, (uint32_t)(local + 1) * kSubBits, slot, adv, nullptr, nullptr, 0u, 0u);
      vin[sub] = start;
      vstate[sub] = (unsigned long long)p | ((unsigned long long)slot << 32);
      advanced[sub] = adv;
      if (!first) *changed = 1;
    }
    __threadfence_block();
    __syncthreads();
  }
}

// exclusive scan of `advanced` inside every stream (one warp per stream, 8 entries per lane per round)
__global__ void huff_scan_kernel(const JpegStream* __restrict__ streams, int n_streams, const uint32_t* __restrict__ advanced,
                                 unsigned long long* __restrict__ slot_start) {
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= n_streams) return;
  const JpegStream st = streams[s];
  unsigned long long run = 0;
  for (int base = 0; base < st.nsub; base += 256) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int i = base + lane * 8 + k;
      v[k] = i < st.nsub ? advanced[st.sub_first + i] : 0u;
    }
    unsigned long long tot = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) tot += v[k];
    unsigned long long x = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    unsigned long long acc = run + x - tot;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int i = base + lane * 8 + k;
      if (i < st.nsub) slot_start[st.sub_first + i] = acc;
      acc += v[k];
    }
    run += __shfl_sync(0xffffffffu, x, 31);
  }
}


__global__ void __launch_bounds__(kHuffThreads)
huff_write_kernel(const JpegImg* __restrict__ imgs, const int* __restrict__ cta_img, const JpegStream* __restrict__ streams,
                  const HuffDev* __restrict__ tabs, const uint32_t* __restrict__ data, const SubState* __restrict__ state,
                  const unsigned long long* __restrict__ slot_start, int16_t* __restrict__ coef_arena, int16_t* __restrict__ dcv) {
  __shared__ HuffSmem hs;
  __shared__ JpegImg im;
  const int img = cta_img[blockIdx.x];
  for (int i = threadIdx.x; i < (int)(sizeof(JpegImg) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(&im)[i] = reinterpret_cast<const uint32_t*>(imgs + img)[i];
  load_huff(hs, tabs, imgs[img].huff_base);
  __syncthreads();
  const int sub = blockIdx.x * kHuffThreads + threadIdx.x;
  if (sub >= im.sub_base + im.nsub) return;
  const int si = find_stream(streams, im.stream_base, im.nstreams, sub);
  const JpegStream st = streams[si];
  const int local = sub - st.sub_first;
  extern __shared__ uint32_t staged[];   // [kHuffThreads][kSubWords]
  stage_sub(staged + threadIdx.x * kSubWords, data, st.bit_off, local);
  uint32_t p = 0, slot = 0;
  if (local) {
    const SubState prev = state[sub - 1];
    p = prev.p;
    slot = prev.slot;
  }
  int16_t* coef[3] = {coef_arena + im.coef_off[0], coef_arena + im.coef_off[1], coef_arena + im.coef_off[2]};
  const uint32_t base_blk = (uint32_t)st.first_mcu * (uint32_t)im.bpm;
  BitWin bw{staged + threadIdx.x * kSubWords, (uint32_t)local * kSubBits, 0ull, 0, 0};
  uint32_t adv = 0;
  huff_run<true>(im, hs, bw, st.bit_off, st.nbits, p, (uint32_t)(local + 1) * kSubBits, slot, adv, coef, dcv,
                 base_blk + (uint32_t)(slot_start[sub] >> 6), base_blk + (uint32_t)st.n_mcu * (uint32_t)im.bpm);
}

// DC differences -> DC values: a scan per (stream, component) over its blocks in MCU order, in three
// steps over 2048-block chunks: chunk sums, an exclusive scan of the sums of each (stream, component),
// then every chunk is scanned again from its offset and written back.
constexpr int kDcChunk = 2048, kDcThreads = 256;   // 8 blocks per thread
struct DcChunk {
  int stream, comp;
  int first, count;      // blocks of this (stream, component), in MCU order
  int group_first;       // index of the first chunk of the same (stream, component)
  int pad;
};
template <bool APPLY>
__global__ void __launch_bounds__(kDcThreads)
dc_chunk_kernel(const JpegImg* __restrict__ imgs, const JpegStream* __restrict__ streams, const DcChunk* __restrict__ chunks,
                int* __restrict__ sums, int16_t* __restrict__ dcv) {
  __shared__ int warp_tot[kDcThreads / 32];
  const DcChunk ck = chunks[blockIdx.x];
  const JpegStream st = streams[ck.stream];
  const JpegImg& im = imgs[st.img];
  // the blocks of (stream, component) in MCU order are contiguous in the compact DC array
  int16_t* coef = dcv + im.dc_off[ck.comp] + (size_t)st.first_mcu * (im.comp_h[ck.comp] * im.comp_v[ck.comp]) + ck.first;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int v[8];
  size_t addr[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const int i = threadIdx.x * 8 + k;
    v[k] = 0;
    addr[k] = 0;
    if (i < ck.count) {
      addr[k] = (size_t)i;
      v[k] = coef[addr[k]];
    }
  }
#pragma unroll
  for (int k = 1; k < 8; k++) v[k] += v[k - 1];
  int x = v[7];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  int before = 0;
  for (int w2 = 0; w2 < warp; w2++) before += warp_tot[w2];
  if (!APPLY) {
    if (threadIdx.x == kDcThreads - 1) sums[blockIdx.x] = before + x;
    return;
  }
  const int off = sums[blockIdx.x] + before + x - v[7];   // sums[] now holds the exclusive prefix of the chunk
#pragma unroll
  for (int k = 0; k < 8; k++)
    if (threadIdx.x * 8 + k < ck.count) coef[addr[k]] = (int16_t)(v[k] + off);
}
// exclusive scan of the chunk sums inside each (stream, component): the first chunk of a group does it
__global__ void dc_offsets_kernel(const DcChunk* __restrict__ chunks, int n_chunks, int* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks || chunks[c].group_first != c) return;
  int run = 0;
  for (int k = c; k < n_chunks && chunks[k].group_first == c; k++) {
    const int t = sums[k];
    sums[k] = run;
    run += t;
  }
}

// jidctint: 8 lanes per block, lane = column in pass 1, row in pass 2 (transpose through shared memory)
__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
__device__ __forceinline__ uint32_t range_limit(int x) {
  int v = x & 1023;
  if (v >= 512) v -= 1024;
  return (uint32_t)min(max(v + 128, 0), 255);
}
__device__ __forceinline__ void idct_1d(const int in[8], int out[8], int shift) {
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * 4433;
  int tmp2 = z1 + z3 * (-15137), tmp3 = z1 + z2 * 6270;
  int tmp0 = (in[0] + in[4]) * 8192, tmp1 = (in[0] - in[4]) * 8192;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * 9633;
  tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
  z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
  z3 += z5; z4 += z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  out[0] = descale(tmp10 + tmp3, shift); out[7] = descale(tmp10 - tmp3, shift);
  out[1] = descale(tmp11 + tmp2, shift); out[6] = descale(tmp11 - tmp2, shift);
  out[2] = descale(tmp12 + tmp1, shift); out[5] = descale(tmp12 - tmp1, shift);
  out[3] = descale(tmp13 + tmp0, shift); out[4] = descale(tmp13 - tmp0, shift);
}

// jidctred.c: reduced-size inverse DCTs on the dequantised coefficients d[64] (natural order) of one block
__device__ __forceinline__ void idct_4x4_store(const int d[64], uint8_t* out, size_t pitch) {
  int ws[4][8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    if (c == 4) continue;   // the second pass never reads column 4
    int tmp0 = d[c] * (1 << 14);
    int tmp2 = d[16 + c] * 15137 + d[48 + c] * (-6270);
    const int tmp10 = tmp0 + tmp2, tmp12 = tmp0 - tmp2;
    const int z1 = d[56 + c], z2 = d[40 + c], z3 = d[24 + c], z4 = d[8 + c];
    tmp0 = z1 * (-1730) + z2 * 11893 + z3 * (-17799) + z4 * 8697;
    tmp2 = z1 * (-4176) + z2 * (-4926) + z3 * 7373 + z4 * 20995;
    ws[0][c] = descale(tmp10 + tmp2, 12); ws[3][c] = descale(tmp10 - tmp2, 12);
    ws[1][c] = descale(tmp12 + tmp0, 12); ws[2][c] = descale(tmp12 - tmp0, 12);
  }
#pragma unroll
  for (int r = 0; r < 4; r++) {
    int tmp0 = ws[r][0] * (1 << 14);
    int tmp2 = ws[r][2] * 15137 + ws[r][6] * (-6270);
    const int tmp10 = tmp0 + tmp2, tmp12 = tmp0 - tmp2;
    const int z1 = ws[r][7], z2 = ws[r][5], z3 = ws[r][3], z4 = ws[r][1];
    tmp0 = z1 * (-1730) + z2 * 11893 + z3 * (-17799) + z4 * 8697;
    tmp2 = z1 * (-4176) + z2 * (-4926) + z3 * 7373 + z4 * 20995;
    uint8_t* o = out + (size_t)r * pitch;
    o[0] = (uint8_t)range_limit(descale(tmp10 + tmp2, 19)); o[3] = (uint8_t)range_limit(descale(tmp10 - tmp2, 19));
    o[1] = (uint8_t)range_limit(descale(tmp12 + tmp0, 19)); o[2] = (uint8_t)range_limit(descale(tmp12 - tmp0, 19));
  }
}
__device__ __forceinline__ void idct_2x2_store(const int d[64], uint8_t* out, size_t pitch) {
  int ws[2][8];
#pragma unroll
  for (int c = 0; c < 8; c++) {
    if (c == 2 || c == 4 || c == 6) continue;
    const int tmp10 = d[c] * (1 << 15);
    const int tmp0 = d[56 + c] * (-5906) + d[40 + c] * 6967 + d[24 + c] * (-10426) + d[8 + c] * 29692;
    ws[0][c] = descale(tmp10 + tmp0, 13);
    ws[1][c] = descale(tmp10 - tmp0, 13);
  }
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int tmp10 = ws[r][0] * (1 << 15);
    const int tmp0 = ws[r][7] * (-5906) + ws[r][5] * 6967 + ws[r][3] * (-10426) + ws[r][1] * 29692;
    out[(size_t)r * pitch] = (uint8_t)range_limit(descale(tmp10 + tmp0, 20));
    out[(size_t)r * pitch + 1] = (uint8_t)range_limit(descale(tmp10 - tmp0, 20));
  }
}

struct IdctJob {                // one component of one image
  unsigned long long coef_off, plane_off;
  int nblocks, bw;              // blocks, blocks per row
  int block_base;               // first global block index of this job
  int qidx;                     // index into the quantisation-table array (64 x u16 each)
  unsigned long long dc_off;    // the component's first entry in the compact DC array
  int ch, cv, mcux, S;          // sampling factors and MCUs per row: block (by, bx) -> MCU-order DC index; IDCT size (8, 4, 2, 1)
};

// One thread per 8x8 block: the block's 128 bytes are read as 8 x 16 bytes (the lanes of a warp read neighbouring
// blocks, so the lines are shared through L1), dequantised, transformed in 64 registers (columns, then rows, as
// jidctint does) and stored as 8 x 8 bytes — neighbouring lanes write neighbouring 8-byte row segments.  The job
// search, the block address and the quantisation row loads are paid once per 64 samples.
constexpr int kIdctThreads = 128;
__global__ void __launch_bounds__(kIdctThreads)
idct_kernel(const IdctJob* __restrict__ jobs, int n_jobs, int total_blocks, const int16_t* __restrict__ coef_arena,
            const uint16_t* __restrict__ qtabs, const int16_t* __restrict__ dcv, uint8_t* __restrict__ plane_arena) {
  const int gb = blockIdx.x * kIdctThreads + threadIdx.x;
  if (gb >= total_blocks) return;
  int ji = 0, hi = n_jobs - 1;   // binary search: a batch has three jobs per image
  while (ji < hi) {
    const int m = (ji + hi + 1) >> 1;
    if (__ldg(&jobs[m].block_base) <= gb) ji = m; else hi = m - 1;
  }
  const IdctJob J = jobs[ji];
  const int b = gb - J.block_base;
  const uint4* src = reinterpret_cast<const uint4*>(coef_arena + J.coef_off + (size_t)b * 64);
  const uint4* qsrc = reinterpret_cast<const uint4*>(qtabs + J.qidx * 64);
  // coefficients and quantisation steps are both stored in zigzag order (the Huffman writer's order: the nonzero
  // ones cluster in a block's first sectors); the permutation to natural order is register renaming here
  constexpr int nat[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                           41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                           30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  int d[64];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const uint4 cw = __ldg(src + r), qw = __ldg(qsrc + r);
    const uint32_t cc[4] = {cw.x, cw.y, cw.z, cw.w}, qq[4] = {qw.x, qw.y, qw.z, qw.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      d[nat[r * 8 + 2 * k]] = (int)(int16_t)(cc[k] & 0xFFFFu) * (int)(qq[k] & 0xFFFFu);
      d[nat[r * 8 + 2 * k + 1]] = (int)(int16_t)(cc[k] >> 16) * (int)(qq[k] >> 16);
    }
  }
  const int by = b / J.bw, bx = b - by * J.bw;
  uint8_t* p = plane_arena + J.plane_off + ((size_t)(by * 8) * J.bw + bx) * 8;
  {   // the predicted DC value comes from the compact array (the arena's slot 0 is never written)
    const int my = by / J.cv, mx = bx / J.ch;
    const size_t di = (size_t)J.dc_off + ((size_t)my * J.mcux + mx) * (J.ch * J.cv) + ((by - my * J.cv) * J.ch + (bx - mx * J.ch));
    d[0] = (int)__ldg(dcv + di) * (int)(__ldg(qtabs + J.qidx * 64));
  }
  if (J.S != 8) {   // reduced-size decode (shrink-on-load): S x S samples per block, plane pitch bw * S
    const size_t pitch = (size_t)J.bw * J.S;
    uint8_t* o = plane_arena + J.plane_off + (size_t)(by * J.S) * pitch + (size_t)bx * J.S;
    if (J.S == 4) idct_4x4_store(d, o, pitch);
    else if (J.S == 2) idct_2x2_store(d, o, pitch);
    else o[0] = (uint8_t)range_limit(descale(d[0], 3));
    return;
  }
#pragma unroll
  for (int c = 0; c < 8; c++) {
    int v[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; r++) v[r] = d[r * 8 + c];
    idct_1d(v, o, 11);
#pragma unroll
    for (int r = 0; r < 8; r++) d[r * 8 + c] = o[r];
  }
#pragma unroll
  for (int r = 0; r < 8; r++) {
    int o[8];
    idct_1d(d + r * 8, o, 18);
    const uint32_t lo = range_limit(o[0]) | (range_limit(o[1]) << 8) | (range_limit(o[2]) << 16) | (range_limit(o[3]) << 24);
    const uint32_t hi2 = range_limit(o[4]) | (range_limit(o[5]) << 8) | (range_limit(o[6]) << 16) | (range_limit(o[7]) << 24);
    *reinterpret_cast<uint2*>(p + (size_t)r * J.bw * 8) = make_uint2(lo, hi2);
  }
}

// fancy upsampling of one chroma sample position + colour conversion, one thread per output pixel
__device__ __forceinline__ int chroma_at(const uint8_t* __restrict__ pl, int pw, int dw, int dh, int hx, int vx, int fancy, int x, int y) {
  if (hx == 1 && vx == 1) return pl[(size_t)y * pw + x];
  // jdsample.c picks the fancy h2v1 / h2v2 upsamplers only for components more than two samples wide; narrower ones
  // (images of 4 pixels or less across) get plain replication, vertically too — and so does everything at scale 1/8
  if (!fancy || (hx == 2 && dw <= 2)) return pl[(size_t)(vx == 2 ? y >> 1 : y) * pw + (hx == 2 ? x >> 1 : x)];
  if (hx == 2 && vx == 1) {
    const uint8_t* in = pl + (size_t)y * pw;
    const int i = x >> 1;
    if (x & 1) {
      if (i == dw - 1 && dw > 1) return in[i];
      return (in[i] * 3 + in[i + 1] + 2) >> 2;
    }
    if (i == 0) return in[0];
    return (in[i] * 3 + in[i - 1] + 1) >> 2;
  }
  const int r = y >> 1;
  const int rn = (y & 1) ? min(r + 1, dh - 1) : max(r - 1, 0);
  const uint8_t* in0 = pl + (size_t)r * pw;
  const uint8_t* in1 = pl + (size_t)rn * pw;
  if (hx == 1) return (in0[x] * 3 + in1[x] + ((y & 1) ? 2 : 1)) >> 2;
  const int i = x >> 1;
  const int thiscol = in0[i] * 3 + in1[i];
  if (x & 1) {
    if (i == dw - 1 && dw > 1) return (thiscol * 4 + 7) >> 4;
    return (thiscol * 3 + in0[i + 1] * 3 + in1[i + 1] + 7) >> 4;
  }
  if (i == 0) return (thiscol * 4 + 8) >> 4;
  return (thiscol * 3 + in0[i - 1] * 3 + in1[i - 1] + 8) >> 4;
}

// is this image the common camera layout (3 components, 4:2:0, not tiny)?  Those take colour420_kernel.
__device__ __forceinline__ bool is_plain_420(const JpegImg& im) {
  return im.ncomp == 3 && im.hmax == 2 && im.vmax == 2 && im.comp_h[1] == 1 && im.comp_v[1] == 1 && im.comp_h[2] == 1 &&
         im.comp_v[2] == 1 && im.w >= 16 && im.comp_S[0] == 8;   // full scale only: a reduced decode takes colour_kernel
}

// 4:2:0: a 8 x 2 patch of output pixels per thread — the two rows share their near chroma row, so each chroma
// plane is read as three rows of (one aligned word + two edge bytes).  Replicating the first / last chroma column
// turns libjpeg's special first- and last-column formulas into the general one ((4c + 8) >> 4 == (3c + c + 8) >> 4).
__device__ __forceinline__ void chroma6(const uint8_t* __restrict__ row, int i0, int dw, int v[6]) {
  if (i0 >= 1 && i0 + 4 <= dw - 1) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(row + i0);   // i0 is a multiple of 4, rows are 8-byte multiples
    v[0] = row[i0 - 1];
    v[1] = (int)(w & 255u); v[2] = (int)((w >> 8) & 255u); v[3] = (int)((w >> 16) & 255u); v[4] = (int)(w >> 24);
    v[5] = row[i0 + 4];
  } else {
#pragma unroll
    for (int k = 0; k < 6; k++) v[k] = row[min(max(i0 - 1 + k, 0), dw - 1)];
  }
}
// YCbCr -> RGB of 8 pixels into 24 bytes of a shared-memory row (the CTA's rows go out as whole 16-byte stores)
__device__ __forceinline__ void ycc_row_stage(uint2 yw, const int (&ub)[8], const int (&ur)[8], uint2* __restrict__ o) {
  uint32_t px[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    // jdcolor.c's Y + ((c * (x - 128) + 32768) >> 16) as one multiply-add chain: Y * 65536 is a multiple of the
    // shift, and the -128 offsets fold into the constant — the additions ride on the FMA pipe, the half-rate ALU
    // pipe keeps the byte extraction, the shifts and the clamps
    const int Y = (int)(((k < 4 ? yw.x : yw.y) >> (8 * (k & 3))) & 0xFF);
    const int y16 = Y * 65536;
    const int rr = (91881 * ur[k] + (y16 + 32768 - 128 * 91881)) >> 16;
    const int g = (-22554 * ub[k] + (-46802 * ur[k] + (y16 + 32768 + 128 * (22554 + 46802)))) >> 16;
    const int b = (116130 * ub[k] + (y16 + 32768 - 128 * 116130)) >> 16;
    px[k] = (uint32_t)__vimin_s32_relu(rr, 255) + (uint32_t)__vimin_s32_relu(g, 255) * 256u + (uint32_t)__vimin_s32_relu(b, 255) * 65536u;
  }
  o[0] = make_uint2(px[0] | (px[1] << 24), (px[1] >> 8) | (px[2] << 16));
  o[1] = make_uint2((px[2] >> 16) | (px[3] << 8), px[4] | (px[5] << 24));
  o[2] = make_uint2((px[5] >> 8) | (px[6] << 16), (px[6] >> 16) | (px[7] << 8));
}
constexpr int kC420RowBytes = 64 * 24;   // one CTA row: 64 threads x 8 pixels x 3 bytes
__global__ void __launch_bounds__(256)
colour420_kernel(const JpegImg* __restrict__ imgs, int n_imgs, const uint8_t* __restrict__ plane_arena, uint8_t* __restrict__ pix_arena) {
  __shared__ __align__(16) uint8_t stage[8][kC420RowBytes];
  const JpegImg& im = imgs[blockIdx.z];
  if (!is_plain_420(im)) return;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int xc = blockIdx.x * 512, x0 = xc + tx * 8, r = blockIdx.y * 4 + ty, y = 2 * r;
  if (xc >= im.w || blockIdx.y * 8 >= im.h) return;   // uniform per CTA
  if (x0 < im.w && y < im.h) {
    const int i0 = x0 >> 1;
    int up[2][2][8];   // [row of the pair][Cb / Cr][pixel]
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const int dw = im.comp_dw[c + 1], dh = im.comp_dh[c + 1], pw = im.comp_bw[c + 1] * 8;
      const uint8_t* base = plane_arena + im.plane_off[c + 1];
      int near_[6], above[6], below[6];
      chroma6(base + (size_t)r * pw, i0, dw, near_);
      chroma6(base + (size_t)max(r - 1, 0) * pw, i0, dw, above);
      chroma6(base + (size_t)min(r + 1, dh - 1) * pw, i0, dw, below);
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int cs[6];
#pragma unroll
        for (int k = 0; k < 6; k++) cs[k] = near_[k] * 3 + (h ? below[k] : above[k]);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          up[h][c][2 * j] = (cs[j + 1] * 3 + cs[j] + 8) >> 4;
          up[h][c][2 * j + 1] = (cs[j + 1] * 3 + cs[j + 2] + 7) >> 4;
        }
      }
    }
    const size_t ypitch = (size_t)im.comp_bw[0] * 8;
    const uint8_t* yp = plane_arena + im.plane_off[0] + (size_t)y * ypitch + x0;   // the Y plane is padded to whole blocks
    ycc_row_stage(*reinterpret_cast<const uint2*>(yp), up[0][0], up[0][1], reinterpret_cast<uint2*>(&stage[2 * ty][tx * 24]));
    ycc_row_stage(*reinterpret_cast<const uint2*>(yp + ((y + 1 < im.h) ? ypitch : 0)), up[1][0], up[1][1],
                  reinterpret_cast<uint2*>(&stage[2 * ty + 1][tx * 24]));
  }
  __syncthreads();
  // 8 rows x 96 16-byte pieces, three per thread; a row ends at its 16-byte pitch (the pad bytes are never read)
  const size_t row_bytes = (size_t)im.out_pitch;
  uint8_t* out0 = pix_arena + im.out_off + (size_t)xc * 3;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int piece = threadIdx.x + k * 256, row = piece / 96, col = (piece - row * 96) * 16;
    const int yy = blockIdx.y * 8 + row;
    if (yy < im.h && (size_t)xc * 3 + col < row_bytes)
      *reinterpret_cast<uint4*>(out0 + (size_t)yy * row_bytes + col) = *reinterpret_cast<const uint4*>(&stage[row][col]);
  }
}

// four output pixels per thread: one word of Y, twelve bytes of RGB
__global__ void __launch_bounds__(256)
colour_kernel(const JpegImg* __restrict__ imgs, int n_imgs, const uint8_t* __restrict__ plane_arena, uint8_t* __restrict__ pix_arena) {
  const int img = blockIdx.z;
  const JpegImg& im = imgs[img];
  if (is_plain_420(im)) return;   // colour420_kernel's
  const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x0 >= im.ow || y >= im.oh) return;
  uint8_t* out = pix_arena + im.out_off + (size_t)y * im.out_pitch;
  const uint8_t* p0 = plane_arena + im.plane_off[0];
  const int pw0 = im.comp_bw[0] * im.comp_S[0];
  const int nvalid = min(4, im.ow - x0);
  uint32_t yw;
  if (im.comp_S[0] == 8) {
    yw = *reinterpret_cast<const uint32_t*>(p0 + (size_t)y * pw0 + x0);   // plane rows are multiples of 8 wide
  } else {   // reduced decode: rows of bw * S samples, any alignment
    yw = 0;
    for (int k = 0; k < nvalid; k++) yw |= (uint32_t)p0[(size_t)y * pw0 + x0 + k] << (8 * k);
  }
  if (im.ncomp == 1) {
    if (nvalid == 4 && (im.out_pitch & 3) == 0)
      *reinterpret_cast<uint32_t*>(out + x0) = yw;
    else
      for (int k = 0; k < nvalid; k++) out[x0 + k] = (uint8_t)(yw >> (8 * k));
    return;
  }
  const uint8_t* p1 = plane_arena + im.plane_off[1];
  const uint8_t* p2 = plane_arena + im.plane_off[2];
  const int pw1 = im.comp_bw[1] * im.comp_S[1], pw2 = im.comp_bw[2] * im.comp_S[2];
  const int hx1 = im.comp_hx[1], vx1 = im.comp_vx[1], hx2 = im.comp_hx[2], vx2 = im.comp_vx[2];
  uint32_t px[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int x = min(x0 + k, im.ow - 1);
    const int Y = (yw >> (8 * k)) & 0xFF;
    const int cb = chroma_at(p1, pw1, im.comp_sdw[1], im.comp_sdh[1], hx1, vx1, im.fancy, x, y);
    const int cr = chroma_at(p2, pw2, im.comp_sdw[2], im.comp_sdh[2], hx2, vx2, im.fancy, x, y);
    const int xb = cb - 128, xr = cr - 128;
    const int r = Y + ((91881 * xr + 32768) >> 16);
    const int g = Y + ((-22554 * xb + 32768 - 46802 * xr) >> 16);
    const int b = Y + ((116130 * xb + 32768) >> 16);
    px[k] = (uint32_t)min(max(r, 0), 255) | ((uint32_t)min(max(g, 0), 255) << 8) | ((uint32_t)min(max(b, 0), 255) << 16);
  }
  if (nvalid == 4) {   // out_pitch is a multiple of 16 and x0 of 4: three aligned words
    uint32_t* o = reinterpret_cast<uint32_t*>(out + 3 * x0);
    o[0] = px[0] | (px[1] << 24);
    o[1] = (px[1] >> 8) | (px[2] << 16);
    o[2] = (px[2] >> 16) | (px[3] << 8);
  } else {
    for (int k = 0; k < nvalid; k++) {
      out[3 * (x0 + k)] = (uint8_t)px[k];
      out[3 * (x0 + k) + 1] = (uint8_t)(px[k] >> 8);
      out[3 * (x0 + k) + 2] = (uint8_t)(px[k] >> 16);
    }
  }
}

}  // namespace irp
