// irp_lib.cu — host side of libirp_b200.so: the C ABI declared in include/irp.h.
//
// Replaces, behind the same contracts, the sharp/libvips calls of
//   server-node/src/services/classifier.js:51-52,107-115,135-143,199-207,296-297
//   server-node/src/middleware/imagePreprocess.js:40-53
// One context per GPU; calls on one context are serialised by a mutex (the Node
// addon / Python mirror give each worker its own context or batch their jobs).
// No CPU fallback: every entry point that touches pixels needs the device.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <functional>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/irp.h"
#include "../../include/irp_spec.h"
#include "grey_tables.inc"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <unordered_map>
#include <immintrin.h>   // the host un-stuffing loop (irp_jpeg_host.inc) copies 32 bytes at a time
#include <deque>
#include <thread>
#include <cuda.h>   // CUtensorMap types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked
#include "irp_classify.cuh"
#include "irp_classify_bulk.cuh"
#include "irp_resize_tma.cuh"
#include "irp_resize_mma.cuh"
#include "irp_jpeg.cuh"
#include "irp_jpeg_prog.cuh"
#include "irp_jpeg_enc.cuh"
#include "irp_resize.cuh"

using namespace irp;

namespace {

thread_local std::string g_create_error;

// grow-only device / pinned-host scratch
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 4096;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct PlanDev {
  int32_t* start;
  int32_t* phase;
  int16_t* coef;
  uint32_t* vpairs;
  uint32_t* hpairs;
  uint32_t* vrows;   // [out][16]  the streaming kernel's per-output-row vertical table
  uint32_t* hcols;   // [out][16]  ... and per-output-column horizontal table
  int n;
  std::vector<int32_t> h_start;   // host copy: tile footprints are sized on the host
  // tensor-core resize (irp_resize_mma.cuh): first source index per output, and the banded coefficient matrices
  // (edge-folded taps split hi / lo, already in the MMA operand layout, de-duplicated) with their index tables
  int32_t* mm_first = nullptr;    // [out]
  int32_t* mm_vmat = nullptr;     // [ceil(out / 128) * 4] as the vertical axis: matrix of each (row block, quarter)
  uint8_t* mm_vmats = nullptr;    // mm_ksv * 2048 bytes each
  int32_t* mm_hmat = nullptr;     // [ceil(out / 32)] as the horizontal axis: matrix of each 32-column strip
  uint8_t* mm_hmats = nullptr;    // kMmChBytes each
  int mm_nt = 0;                  // taps per output after folding
  int mm_rows = 0, mm_ksv = 0;    // as the vertical axis: source rows a 128-row tile touches, 32-row K steps per quarter
  bool mm_v_ok = false, mm_h_ok = false;
};

}  // namespace

struct irp_request {
  irp_image_desc img;
  irp_jpeg_desc jpeg;      // used instead of img when is_jpeg_file
  irp_jpeg_out* enc = nullptr;   // file in -> scores + preprocessed FILE out (irp_submit_transcode)
  int quality = 85;
  bool is_jpeg_file = false;
  irp_result* result;
  irp_out_desc* out;
  int status = 0;
  bool done = false;
  std::string err;
};

struct irp_ctx {
  int device = 0;
  int sm_count = 0;
  irp_opts opts{};
  cudaStream_t own_stream = nullptr, stream = nullptr;
  std::mutex mu;
  std::string err;
  ClassifyTables* d_tables = nullptr;
  DevBuf d_desc, d_acc, d_stage_in, d_stage_out, d_orient, d_jobs, d_tmaps, d_rtjobs, d_rtmaps;
  PinBuf h_desc, h_acc, h_jobs, h_tmaps, h_rtjobs, h_rtmaps;
  size_t smem_optin_full = 0;     // the device's opt-in shared memory per block
  DevBuf d_jdata, d_jmeta, d_jstate, d_jcoef, d_jplane, d_jpix;   // device JPEG decode (irp_jpeg.cuh)
  PinBuf h_jdata, h_jmeta;
  DevBuf d_emeta, d_eblk, d_ebits, d_eout, d_epix, d_ehuff;        // device JPEG encode (irp_jpeg_enc.cuh)
  PinBuf h_emeta, h_ehuff;
  int jpeg_sweeps = 0;            // synchronisation sweeps of the last JPEG batch
  void* encode_tiled = nullptr;   // cuTensorMapEncodeTiled
  std::vector<void*> plan_chunks;
  size_t plan_total = 0;   // bytes of all plan chunks (the cache is flushed whole past IRP_PLAN_CACHE_MB, default 2048)
  size_t plan_used = 0, plan_cap = 0;
  std::map<std::tuple<int, int, double>, PlanDev> plans;
  cudaEvent_t ev[6]{};
  bool jpeg_allow_scale = false;    // set by irp_decode_jpeg_batch around its decode: the other entry points classify, and that needs the full picture
  cudaEvent_t ev_block = nullptr;   // cudaEventBlockingSync: wait_stream() sleeps on it under IRP_BLOCKING_SYNC=1
  irp_timing timing{};
  int occ_classify[5]{};  // CTAs per SM for C = 1, 3, 4
  cudaStream_t copy_in_stream = nullptr, copy_out_stream = nullptr;  // H2D / D2H of pipelined host batches
  std::vector<cudaEvent_t> sync_events, timing_events;
  size_t chunk_bytes = 96u << 20;    // pixels per pipeline chunk of a host-resident batch (tools/pcie_probe.py: 64-128 MiB is the flat optimum)
  size_t smem_optin = 0;             // opt-in dynamic shared memory limit of the device
  uint32_t smem_base = 0; // .shared address where dynamic shared memory starts (probed once)
  int* d_error_flag = nullptr;
  // concurrent single-image requests (irp_submit / irp_wait)
  std::thread dispatcher;
  std::mutex qmu;
  std::condition_variable qcv, done_cv;
  std::deque<struct irp_request*> queue;
  bool stop = false, dispatcher_started = false;
  bool rtma_ok = true;    // IRP_NO_RTMA=1 keeps the generic resize kernel (A/B runs)
  bool rmma_ok = false;   // the tensor-core resize kernel (IRP_NO_RMMA=1 turns it off)
  DevBuf d_mmjobs, d_mmmaps, d_mmctr;
  PinBuf h_mmjobs, h_mmmaps;
  bool bulk_ok = false;   // the streaming classify kernel's shared-memory map fits this device
  int* h_error_flag = nullptr;
  // compressed-input batches are cut into lanes: child contexts (own streams and scratch) driven by their own host
  // threads, so one lane's marker scan / un-stuffing / uploads / size read-backs run under the other lanes' kernels
  std::vector<irp_ctx*> lanes;
  // ICC profiles an encode call can name (IRP_JPEG_ICC(id) in `quality`): slot 0 = the context default
  // (irp_set_output_icc), slot IRP_ICC_SRGB = the generated sRGB profile, then irp_register_icc's.  Entries are
  // immutable once published; a call holds a shared_ptr to the one it encodes with.  Lanes resolve through `parent`.
  std::mutex icc_mu;
  std::vector<std::shared_ptr<const std::vector<uint8_t>>> icc_profiles;
  irp_ctx* parent = nullptr;
  std::mutex lanes_mu;
  bool is_lane = false;
  std::mutex pin_mu;
  std::vector<void*> pinned;      // irp_host_alloc_pinned allocations still alive: freed with the context
};

namespace {

int fail(irp_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->err = buf;
  else
    g_create_error = buf;
  return code;
}

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) return fail(ctx, IRP_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// This process's share of the host cores: IRP_HOST_THREADS, else the cores divided by the ranks torchrun started on this
// box (LOCAL_WORLD_SIZE) — eight ranks each helping themselves to all 32 cores only thrash.
inline int host_core_share() {
  static const int share = [] {
    if (const char* e = getenv("IRP_HOST_THREADS")) return std::max(1, atoi(e));
    const char* lw = getenv("LOCAL_WORLD_SIZE");
    const int ranks = lw ? std::max(1, atoi(lw)) : 1;
    return std::max(2, (int)std::thread::hardware_concurrency() / ranks);
  }();
  return share;
}
// IRP_BLOCKING_SYNC=1: host threads that wait for the GPU sleep on a blocking event instead of spinning in
// cudaStreamSynchronize.  Off by default: measured with eight ranks on a 32-vCPU box it frees cores for the other ranks'
// un-stuffing workers (files route 44.0 -> 41.8 ms per step, together with four lanes instead of eight) but the wake-up
// latency after every call costs the device-resident loop far more (2.80 -> 3.27 ms per step).
inline bool host_waits_block() {
  static const bool b = [] {
    const char* e = getenv("IRP_BLOCKING_SYNC");
    return e && atoi(e) != 0;
  }();
  return b;
}

// ---- P3 geometry: imagePreprocess.js:7-22 + sharp ResolveShrink (fit inside, no enlargement) ----
void orient_dims(int w, int h, int o, int* ow, int* oh) {
  if (o >= 5 && o <= 8) {
    *ow = h;
    *oh = w;
  } else {
    *ow = w;
    *oh = h;
  }
}
void fit_inside(int wo, int ho, int tw, int th, int* ow, int* oh, double* shrink) {
  double hs = (double)wo / tw, vs = (double)ho / th;
  double f = std::max(hs, vs);
  f = std::max(f, 1.0);
  f = std::min(f, (double)wo);
  f = std::min(f, (double)ho);
  *shrink = f;
  *ow = std::max(1, (int)std::round((double)wo / f));
  *oh = std::max(1, (int)std::round((double)ho / f));
}
void preprocess_dims(int w, int h, int o, int* ow, int* oh, double* shrink) {
  int wo, ho;
  orient_dims(w, h, o, &wo, &ho);
  if (w > IRP_MAX_DIMENSION || h > IRP_MAX_DIMENSION) {
    double scale = (double)IRP_MAX_DIMENSION / std::max(w, h);
    int tw = std::max(1, (int)std::floor(w * scale + 0.5)), th = std::max(1, (int)std::floor(h * scale + 0.5));
    fit_inside(wo, ho, tw, th, ow, oh, shrink);
  } else {
    *ow = wo;
    *oh = ho;
    *shrink = 1.0;
  }
}
void fusion_dims(int w, int h, int o, int* ow, int* oh, int* ox, int* oy, double* shrink) {
  int wo, ho;
  orient_dims(w, h, o, &wo, &ho);
  fit_inside(wo, ho, IRP_FUSION_CANVAS, IRP_FUSION_CANVAS, ow, oh, shrink);
  *ox = (IRP_FUSION_CANVAS - *ow) / 2;
  *oy = (IRP_FUSION_CANVAS - *oh) / 2;
}

// ---- P3 coefficients: libvips reduce{v,h} lanczos3 masks, 65 phases, 12-bit fixed point ----
int reduce_points(double shrink) { return 2 * (int)std::nearbyint(IRP_LANCZOS_A * shrink) + 1; }

void lanczos_mask(double* c, int n, double shrink, double x) {
  const double a = IRP_LANCZOS_A, pi = 3.14159265358979323846;
  const double half = x + n / 2 - 1;
  double sum = 0;
  for (int i = 0; i < n; i++) {
    double xp = (i - half) / shrink, l;
    if (xp == 0.0)
      l = 1.0;
    else if (xp < -a || xp > a)
      l = 0.0;
    else
      l = a * std::sin(pi * xp) * std::sin(pi * xp / a) / (pi * pi * xp * xp);
    c[i] = l;
    sum += l;
  }
  for (int i = 0; i < n; i++) c[i] /= sum;
}
void to_fixed_point(const double* in, int16_t* out, int n, int scale) {
  double fsum = 0;
  for (int i = 0; i < n; i++) fsum += in[i];
  int target = (int)std::nearbyint(fsum * scale), sum;
  double high = scale + (n + 1) / 2, low = scale - (n + 1) / 2, guess;
  do {
    guess = (high + low) / 2.0;
    sum = 0;
    for (int i = 0; i < n; i++) {
      out[i] = (int16_t)std::nearbyint(in[i] * guess);
      sum += out[i];
    }
    if (sum == target) break;
    if (sum < target) low = guess;
    if (sum > target) high = guess;
  } while (high - low > 0.01);
  if (sum != target) {
    int each = (target - sum) / n, extra = (target - sum) % n;
    int dir = extra > 0 ? 1 : -1, cnt = std::abs(extra);
    for (int i = 0; i < n; i++) out[i] += each;
    for (int i = 0; i < cnt; i++) out[i] += dir;
  }
}
struct HostPlan {
  int n;
  std::vector<int32_t> start, phase;
  std::vector<int16_t> coef;  // [65][kCoefStride]
};
bool build_plan(int in_size, int out_size, double shrink, int coef_mode, int reduce_mode, HostPlan* hp) {
  int n = reduce_points(shrink);
  if (n > IRP_MAX_TAPS) return false;
  hp->n = n;
  hp->coef.assign((size_t)(IRP_PHASES + 1) * kCoefStride, 0);
  double mask[IRP_MAX_TAPS];
  for (int t = 0; t <= IRP_PHASES; t++) {
    lanczos_mask(mask, n, shrink, (double)t / IRP_PHASES);
    int16_t* ci = hp->coef.data() + (size_t)t * kCoefStride;
    if (reduce_mode == IRP_REDUCE_VECTOR_2_6) {   // 6 fractional bits, kept as multiples of 64: (sum 64 c p + 2048) >> 12 == (sum c p + 32) >> 6
      to_fixed_point(mask, ci, n, 64);
      for (int i = 0; i < n; i++) ci[i] = (int16_t)(ci[i] * 64);
    } else if (coef_mode == IRP_COEF_TRUNCATE)
      for (int i = 0; i < n; i++) ci[i] = (int16_t)(mask[i] * (1 << IRP_INTERP_SHIFT));
    else
      to_fixed_point(mask, ci, n, 1 << IRP_INTERP_SHIFT);
  }
  hp->start.resize(out_size);
  hp->phase.resize(out_size);
  double extra = out_size * shrink - in_size;
  for (int o = 0; o < out_size; o++) {
    double c = (o + 0.5) * shrink - 0.5 - extra / 2.0;
    int p = (int)std::floor(c);
    int sy = (int)std::floor(c * IRP_PHASES * 2);
    hp->phase[o] = ((sy & (IRP_PHASES * 2 - 1)) + 1) >> 1;
    hp->start[o] = p - (n / 2 - 1);
  }
  return true;
}

// wait for a stream of this context (the caller holds ctx->mu, so the event is not shared)
cudaError_t wait_stream(irp_ctx* ctx, cudaStream_t st) {
  if (!host_waits_block() || !ctx->ev_block) return cudaStreamSynchronize(st);
  const cudaError_t e = cudaEventRecord(ctx->ev_block, st);
  return e != cudaSuccess ? e : cudaEventSynchronize(ctx->ev_block);
}

int plan_alloc(irp_ctx* ctx, size_t bytes, void** out) {
  bytes = round_up(bytes, 256);
  if (ctx->plan_chunks.empty() || ctx->plan_used + bytes > ctx->plan_cap) {
    size_t cap = std::max<size_t>(bytes, 4u << 20);
    void* p = nullptr;
    CK(cudaMalloc(&p, cap));
    ctx->plan_chunks.push_back(p);
    ctx->plan_total += cap;
    ctx->plan_used = 0;
    ctx->plan_cap = cap;
  }
  *out = (char*)ctx->plan_chunks.back() + ctx->plan_used;
  ctx->plan_used += bytes;
  return IRP_OK;
}

// A plan's tables are collected in one host blob and reach the device as ONE allocation and ONE copy: a dozen
// synchronous small copies were a third of the 1-3 ms a new geometry cost.
struct PlanBlob {
  std::vector<uint8_t> bytes;
  size_t add(const void* data, size_t n) {
    const size_t off = round_up(bytes.size(), 256);
    bytes.resize(off + n);
    if (n) memcpy(bytes.data() + off, data, n);
    return off;
  }
};

// Tables of the tensor-core resize (irp_resize_mma.cuh) for one axis: per output its taps with the replicate edge
// folded in (a tap that falls outside the image is added to the edge tap), split c = 128 * hi + lo, and the first
// source index they apply to.  From those, the banded coefficient matrices the kernel multiplies with, laid out as
// the MMA's K-major N x K operand (16-byte rows of eight-row groups: byte (n, k) at (n >> 3) * 128 + (n & 7) * 16 +
// (k >> 4) * 1024 + (k & 15); n = hi | lo half * 32 + output within the group of 32): one per (128-row block,
// quarter) for the axis as the vertical one, one per 32-column strip as the horizontal one; identical matrices are
// stored once (a periodic geometry such as 4000 -> 2048 needs a handful).
struct MmOffsets { size_t first = 0, vmat = 0, vmats = 0, hmat = 0, hmats = 0; };
int build_mm_plan(const HostPlan& hp, int in_size, PlanDev* pd, PlanBlob* blob, MmOffsets* mo) {
  const int no = (int)hp.start.size(), n = hp.n;
  if (n > kMmMaxTaps) return IRP_OK;
  std::vector<int8_t> tab((size_t)no * 64, 0);
  std::vector<int32_t> first(no);
  int nt_max = 1;
  for (int o = 0; o < no; o++) {
    const int s0 = hp.start[o];
    const int16_t* ci = hp.coef.data() + (size_t)hp.phase[o] * kCoefStride;
    const int lo_idx = std::min(std::max(s0, 0), in_size - 1), hi_idx = std::min(std::max(s0 + n - 1, 0), in_size - 1);
    int acc[kMmMaxTaps] = {0};
    for (int t = 0; t < n; t++) acc[std::min(std::max(s0 + t, 0), in_size - 1) - lo_idx] += ci[t];
    const int nt = hi_idx - lo_idx + 1;
    for (int j = 0; j < nt; j++) {
      const int hi = (acc[j] + 64) >> 7, lo = acc[j] - 128 * hi;
      if (hi < -128 || hi > 127) return IRP_OK;   // cannot happen for lanczos masks; the ALU kernels serve it
      tab[(size_t)o * 64 + j] = (int8_t)hi;
      tab[(size_t)o * 64 + 32 + j] = (int8_t)lo;
    }
    first[o] = lo_idx;
    nt_max = std::max(nt_max, nt);
  }
  for (int o = 1; o < no; o++)
    if (first[o] < first[o - 1]) return IRP_OK;   // windows must not move backwards
  pd->mm_nt = nt_max;
  auto put = [&](std::vector<uint8_t>& m, int nrow, int o, int koff) {   // output o's taps into row nrow (and 32 + nrow) from K index koff
    for (int hl = 0; hl < 2; hl++) {
      const int nn = hl * 32 + nrow;
      for (int j = 0; j < nt_max; j++) {
        const int k = koff + j;
        m[(size_t)(nn >> 3) * 128 + (nn & 7) * 16 + (size_t)(k >> 4) * 1024 + (k & 15)] = (uint8_t)tab[(size_t)o * 64 + hl * 32 + j];
      }
    }
  };
  auto dedupe = [](std::unordered_multimap<uint64_t, int>& seen, std::vector<uint8_t>& store, const std::vector<uint8_t>& m) {
    uint64_t h = 1469598103934665603ull;
    const uint64_t* w = reinterpret_cast<const uint64_t*>(m.data());
    for (size_t k = 0; k < m.size() / 8; k++) h = (h ^ w[k]) * 1099511628211ull;
    auto range = seen.equal_range(h);
    for (auto it = range.first; it != range.second; ++it)
      if (!memcmp(store.data() + (size_t)it->second * m.size(), m.data(), m.size())) return it->second;
    const int idx = (int)(store.size() / m.size());
    store.insert(store.end(), m.begin(), m.end());
    seen.emplace(h, idx);
    return idx;
  };
  // ---- as the vertical axis: the tile's box starts at first[o0]; each quarter's window at an 8-row boundary of the box
  const int tiles_y = (no + kMmTR - 1) / kMmTR;
  auto ws_of = [&](int o0, int q) { return (first[std::min(o0 + 32 * q, no - 1)] - first[o0]) & ~7; };
  int rows = 16, ksv = 1;
  for (int rb = 0; rb < tiles_y; rb++) {
    const int o0 = rb * kMmTR;
    for (int q = 0; q < 4; q++)
      for (int o = o0 + 32 * q; o < std::min(no, o0 + 32 * q + 32); o++) {
        const int koff = first[o] - first[o0] - ws_of(o0, q);
        ksv = std::max(ksv, (koff + nt_max + 31) / 32);
        rows = std::max(rows, first[o] + nt_max - first[o0]);
      }
  }
  pd->mm_rows = (int)round_up((size_t)rows, 16);
  pd->mm_ksv = ksv;
  pd->mm_v_ok = ksv <= kMmMaxKsv;
  std::vector<int32_t> vmat((size_t)tiles_y * 4, 0);
  std::vector<uint8_t> vmats;
  if (pd->mm_v_ok) {
    std::unordered_multimap<uint64_t, int> seen;   // content hash -> matrix index (compared byte for byte on a hit)
    std::vector<uint8_t> m((size_t)ksv * 2048);
    for (int rb = 0; rb < tiles_y; rb++) {
      const int o0 = rb * kMmTR;
      for (int q = 0; q < 4; q++) {
        std::fill(m.begin(), m.end(), 0);
        for (int o = o0 + 32 * q; o < std::min(no, o0 + 32 * q + 32); o++) put(m, o - o0 - 32 * q, o, first[o] - first[o0] - ws_of(o0, q));
        vmat[(size_t)rb * 4 + q] = dedupe(seen, vmats, m);
      }
    }
  }
  // ---- as the horizontal axis: 32-column strips over 256 source bytes starting at a 16-byte boundary
  const int tiles_x = (no + kMmTC - 1) / kMmTC;
  bool h_ok = true;
  for (int b = 0; b < tiles_x && h_ok; b++) {
    const int o0 = b * kMmTC, o1 = std::min(no, o0 + kMmTC) - 1;
    const int bx0 = (3 * first[o0]) & ~15, p0 = bx0 / 3, plast = first[o1] + nt_max - 1;
    if (3 * plast + 2 > bx0 + kMmXB - 1 || plast - p0 >= kMmKH - 1) h_ok = false;   // K index 95 is the rounding column
  }
  pd->mm_h_ok = h_ok;
  std::vector<int32_t> hmat(tiles_x, 0);
  std::vector<uint8_t> hmats;
  if (h_ok) {
    std::unordered_multimap<uint64_t, int> seen;
    std::vector<uint8_t> m(kMmChBytes);
    for (int b = 0; b < tiles_x; b++) {
      const int o0 = b * kMmTC, p0 = ((3 * first[o0]) & ~15) / 3;
      std::fill(m.begin(), m.end(), 0);
      for (int o = o0; o < std::min(no, o0 + kMmTC); o++) put(m, o - o0, o, first[o] - p0);
      // reduceh's rounding constant: the intermediate's pixel 95 is a constant 1, every hi row holds 16 there (16 * 128 = 2048)
      for (int nn = 0; nn < kMmTC; nn++) m[(size_t)(nn >> 3) * 128 + (nn & 7) * 16 + (size_t)((kMmKH - 1) >> 4) * 1024 + ((kMmKH - 1) & 15)] = 16;
      hmat[b] = dedupe(seen, hmats, m);
    }
  }
  mo->first = blob->add(first.data(), first.size() * 4);
  if (pd->mm_v_ok) {
    mo->vmat = blob->add(vmat.data(), vmat.size() * 4);
    mo->vmats = blob->add(vmats.data(), vmats.size());
  }
  if (h_ok) {
    mo->hmat = blob->add(hmat.data(), hmat.size() * 4);
    mo->hmats = blob->add(hmats.data(), hmats.size());
  }
  return IRP_OK;
}

int get_plan(irp_ctx* ctx, int in_size, int out_size, double shrink, AxisPlan* ap, const PlanDev** pdev = nullptr) {
  const bool identity = in_size == out_size;
  auto key = std::make_tuple(in_size, out_size, identity ? 1.0 : shrink);
  auto it = ctx->plans.find(key);
  if (it == ctx->plans.end()) {
    HostPlan hp;
    if (identity) {  // one tap of weight 1.0: (4096 * p + 2048) >> 12 == p
      hp.n = 1;
      hp.coef.assign((size_t)(IRP_PHASES + 1) * kCoefStride, 0);
      for (int t = 0; t <= IRP_PHASES; t++) hp.coef[(size_t)t * kCoefStride] = 1 << IRP_INTERP_SHIFT;
      hp.start.resize(out_size);
      hp.phase.assign(out_size, 0);
      for (int o = 0; o < out_size; o++) hp.start[o] = o;
    } else if (!build_plan(in_size, out_size, shrink, ctx->opts.coef_mode, ctx->opts.reduce_mode, &hp)) {
      return fail(ctx, IRP_ERR_UNSUPPORTED, "shrink factor %.4f needs more than %d taps", shrink, IRP_MAX_TAPS);
    }
    PlanDev pd;
    pd.n = hp.n;
    int rc;
    PlanBlob blob;
    // coefficient pairs pre-shifted for the kernels: pair p of shift s = (c[2p - s], c[2p + 1 - s])
    std::vector<uint32_t> vp((size_t)(IRP_PHASES + 1) * 2 * 16, 0), hq((size_t)(IRP_PHASES + 1) * 4 * 16, 0);
    for (int t = 0; t <= IRP_PHASES; t++) {
      const int16_t* ci = hp.coef.data() + (size_t)t * kCoefStride;
      auto tap = [&](int i) -> uint32_t { return (i >= 0 && i < hp.n) ? (uint16_t)ci[i] : 0u; };
      for (int sft = 0; sft < 2; sft++)
        for (int k = 0; k < kMaxPairs; k++) vp[((size_t)t * 2 + sft) * 16 + k] = tap(2 * k - sft) | (tap(2 * k + 1 - sft) << 16);
      for (int sft = 0; sft < 4; sft++)
        for (int k = 0; k < kMaxHPairs; k++) hq[((size_t)t * 4 + sft) * 16 + k] = tap(2 * k - sft) | (tap(2 * k + 1 - sft) << 16);
    }
    // one 16-word row per output row / column: its coefficient pairs and where its window starts
    const size_t no = hp.start.size();
    std::vector<uint32_t> vr(no * 16, 0), hc(no * kHTabWords, 0);
    for (size_t o = 0; o < no; o++) {
      const int s0 = hp.start[o], ph = hp.phase[o];
      for (int k = 0; k < kMaxPairs; k++) vr[o * 16 + k] = vp[((size_t)ph * 2 + (s0 & 1)) * 16 + k];
      vr[o * 16 + 15] = (uint32_t)(s0 >> 1);
      for (int k = 0; k < kMaxHPairs; k++) hc[o * kHTabWords + k] = hq[((size_t)ph * 4 + (s0 & 3)) * 16 + k];
      hc[o * kHTabWords + 15] = (uint32_t)(s0 >> 2);
    }
    const size_t o_start = blob.add(hp.start.data(), hp.start.size() * 4), o_phase = blob.add(hp.phase.data(), hp.phase.size() * 4),
                 o_coef = blob.add(hp.coef.data(), hp.coef.size() * 2), o_vp = blob.add(vp.data(), vp.size() * 4),
                 o_hq = blob.add(hq.data(), hq.size() * 4), o_vr = blob.add(vr.data(), vr.size() * 4), o_hc = blob.add(hc.data(), hc.size() * 4);
    pd.h_start = hp.start;
    MmOffsets mo;
    if (!identity && (rc = build_mm_plan(hp, in_size, &pd, &blob, &mo))) return rc;
    void* base;
    if ((rc = plan_alloc(ctx, blob.bytes.size(), &base))) return rc;
    CK(cudaMemcpy(base, blob.bytes.data(), blob.bytes.size(), cudaMemcpyHostToDevice));   // synchronous: built once per geometry and cached
    uint8_t* d = (uint8_t*)base;
    pd.start = (int32_t*)(d + o_start);
    pd.phase = (int32_t*)(d + o_phase);
    pd.coef = (int16_t*)(d + o_coef);
    pd.vpairs = (uint32_t*)(d + o_vp);
    pd.hpairs = (uint32_t*)(d + o_hq);
    pd.vrows = (uint32_t*)(d + o_vr);
    pd.hcols = (uint32_t*)(d + o_hc);
    if (!identity) {
      pd.mm_first = (int32_t*)(d + mo.first);
      if (pd.mm_v_ok) {
        pd.mm_vmat = (int32_t*)(d + mo.vmat);
        pd.mm_vmats = d + mo.vmats;
      }
      if (pd.mm_h_ok) {
        pd.mm_hmat = (int32_t*)(d + mo.hmat);
        pd.mm_hmats = d + mo.hmats;
      }
    }
    it = ctx->plans.emplace(key, pd).first;
  }
  *ap = AxisPlan{it->second.start, it->second.phase, it->second.coef, it->second.vpairs, it->second.hpairs, it->second.n, 0};
  if (pdev) *pdev = &it->second;
  return IRP_OK;
}

// ---- JS formulas on exact integer moments ----
double pop_variance(uint64_t s, uint64_t q, uint64_t n) {
  // (n*q - s^2) / n^2, numerator exact in 128 bits
  unsigned __int128 num = (unsigned __int128)n * q - (unsigned __int128)s * s;
  return (double)num / ((double)n * (double)n);
}
double jsmin(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }
double jsmax(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }

int validate_desc(irp_ctx* ctx, const irp_image_desc& d, int i) {
  if (!d.pixels) return fail(ctx, IRP_ERR_BAD_ARG, "image %d: null pixels", i);
  if (d.width <= 0 || d.height <= 0) return fail(ctx, IRP_ERR_BAD_ARG, "image %d: bad dims %dx%d", i, d.width, d.height);
  if (d.channels != 1 && d.channels != 3 && d.channels != 4)
    return fail(ctx, IRP_ERR_UNSUPPORTED, "image %d: %d channels unsupported (1, 3 or 4)", i, d.channels);
  if (d.pitch < (size_t)d.width * d.channels) return fail(ctx, IRP_ERR_BAD_ARG, "image %d: pitch < width*channels", i);
  return IRP_OK;
}

struct Staged {  // where each input image lives on the device for this call
  const uint8_t* px;
  size_t pitch;
  size_t bytes;   // bytes that cross PCIe for this image (0 if device-resident)
};

struct Geo {      // per-image preprocess geometry
  int wo, ho, dw, dh, dc, ox, oy, o;
  double f;
  size_t orient_off, out_off;
  // vips_resize's integer pre-shrink (axes shrinking 4x or more): factors, what the lanczos passes then see
  int kh = 1, kv = 1, rw = 0, rh = 0;
  double fh = 1.0, fv = 1.0;
  size_t box_off = 0;
};

struct OutPlan {  // per image: where the kernel writes, and how the result gets to the caller
  uint8_t* dev;
  size_t dev_pitch;
  bool via_stage;
};

cudaError_t get_event(irp_ctx* ctx, size_t idx, cudaEvent_t* ev, bool timing) {
  std::vector<cudaEvent_t>& pool = timing ? ctx->timing_events : ctx->sync_events;
  while (pool.size() <= idx) {
    cudaEvent_t e;
    cudaError_t rc = timing ? cudaEventCreate(&e) : cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (rc != cudaSuccess) return rc;
    pool.push_back(e);
  }
  *ev = pool[idx];
  return cudaSuccess;
}

// decide where every input lives on the device; host inputs get a slot in the staging arena
int plan_inputs(irp_ctx* ctx, const irp_image_desc* imgs, int n, std::vector<Staged>* st) {
  st->resize(n);
  size_t total = 0;
  std::vector<size_t> off(n, 0);
  for (int i = 0; i < n; i++) {
    if (!imgs[i].pixels || imgs[i].on_device) continue;
    size_t pitch = round_up((size_t)imgs[i].width * imgs[i].channels, 16);
    off[i] = total;
    total += round_up(pitch * imgs[i].height, 256);
  }
  if (total) CK(ctx->d_stage_in.reserve(total + 256));
  for (int i = 0; i < n; i++) {
    const irp_image_desc& d = imgs[i];
    if (!d.pixels) {
      (*st)[i] = Staged{nullptr, 0, 0};
    } else if (d.on_device) {
      (*st)[i] = Staged{d.pixels, d.pitch, 0};
    } else {
      size_t pitch = round_up((size_t)d.width * d.channels, 16);
      (*st)[i] = Staged{(uint8_t*)ctx->d_stage_in.p + off[i], pitch, (size_t)d.width * d.channels * d.height};
    }
  }
  return IRP_OK;
}

// rows -> rows; one linear copy when both sides are tight (a 2-D descriptor per image costs the copy engines more than
// a linear one, which matters when eight ranks share one host)
static cudaError_t copy_rows_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t row_bytes, size_t rows, cudaMemcpyKind kind,
                                   cudaStream_t stream) {
  if (dpitch == row_bytes && spitch == row_bytes) return cudaMemcpyAsync(dst, src, row_bytes * rows, kind, stream);
  return cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, kind, stream);
}

int copy_inputs(irp_ctx* ctx, const irp_image_desc* imgs, const std::vector<Staged>& st, int b, int e, cudaStream_t stream) {
  for (int i = b; i < e; i++) {
    const irp_image_desc& d = imgs[i];
    if (!d.pixels || d.on_device) continue;
    CK(copy_rows_async((void*)st[i].px, st[i].pitch, d.pixels, d.pitch, (size_t)d.width * d.channels, d.height,
                       cudaMemcpyHostToDevice, stream));
  }
  return IRP_OK;
}

template <int C>
int launch_classify(irp_ctx* ctx, const ImgDev* d_imgs, int n, int total_tiles, unsigned long long* d_acc,
                    uint32_t* d_hist) {
  if (!n) return IRP_OK;
  const SmemMap map = make_smem_map(ctx->smem_base, C);
  const size_t smem = map.end - ctx->smem_base;
  int& occ = ctx->occ_classify[C];
  if (!occ) {
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, classify_kernel<C>, kClassifyThreads, smem));
    if (occ < 1) return fail(ctx, IRP_ERR_CUDA, "classify kernel does not fit on an SM");
  }
  int grid = std::min((total_tiles + kGroups - 1) / kGroups, ctx->sm_count * occ);
  classify_kernel<C><<<grid, kClassifyThreads, smem, ctx->stream>>>(d_imgs, n, total_tiles, ctx->d_tables, d_acc, d_hist,
                                                                    (uint32_t)smem, ctx->d_error_flag,
                                                                    ctx->opts.blur_mode == IRP_BLUR_VECTOR ? blur_vector() : blur_exact());
  CK(cudaGetLastError());
  ctx->timing.kernel_launches++;
  return IRP_OK;
}

// the streaming kernel: 3-channel images with 16-byte aligned base and pitch (irp_classify_bulk.cuh)
int launch_classify_bulk(irp_ctx* ctx, const ImgDev* d_imgs, const TmaDesc* d_tmaps, int n, int total_tiles, unsigned long long* d_acc,
                         uint32_t* d_hist) {
  if (!n) return IRP_OK;
  const BulkMap map = make_bulk_map(ctx->smem_base);
  const size_t smem = map.end - ctx->smem_base;
  int grid = std::min((total_tiles + kBGroups - 1) / kBGroups, ctx->sm_count);
  classify_bulk_kernel<<<grid, kBThreads, smem, ctx->stream>>>(d_imgs, d_tmaps, n, total_tiles, ctx->d_tables, d_acc, d_hist,
                                                              (uint32_t)smem, ctx->d_error_flag, MulConsts{4u});
  CK(cudaGetLastError());
  ctx->timing.kernel_launches++;
  return IRP_OK;
}

// the image as the TMA unit sees it: rows of u16 elements, box = one raw tile (208 x 34)
int encode_image_tmap(irp_ctx* ctx, const ImgDev& d, TmaDesc* out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                               CUtensorMapFloatOOBfill);
  static_assert(sizeof(TmaDesc) == sizeof(CUtensorMap), "TmaDesc must mirror CUtensorMap");
  const cuuint64_t dims[2] = {(cuuint64_t)(d.w * 3 + 1) / 2, (cuuint64_t)d.h};
  const cuuint64_t strides[1] = {(cuuint64_t)d.pitch};
  const cuuint32_t box[2] = {(cuuint32_t)kRawPitch / 2, (cuuint32_t)kRows}, estr[2] = {1, 1};
  CUresult r = ((EncodeFn)ctx->encode_tiled)(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void*)d.px, dims,
                                             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, IRP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %dx%d image, pitch %llu", (int)r, d.w, d.h, d.pitch);
  return IRP_OK;
}

inline size_t acc_bytes_for(int n) { return round_up(sizeof(unsigned long long) * ACC_COUNT * n, 256); }

// classify images [b, e): their descriptors occupy slots [b, e) of the batch-wide arrays, grouped by
// kernel (1 channel, 3 channels generic, 4 channels, 3 channels streaming); each group is one launch
// over all of its tiles
int classify_range(irp_ctx* ctx, const irp_image_desc* imgs, const std::vector<Staged>& st, int n, int b, int e) {
  ImgDev* h_imgs = (ImgDev*)ctx->h_desc.p;
  TmaDesc* h_tmaps = (TmaDesc*)(((uintptr_t)ctx->h_tmaps.p + 63) & ~(uintptr_t)63);
  TmaDesc* d_tmaps = (TmaDesc*)(((uintptr_t)ctx->d_tmaps.p + 63) & ~(uintptr_t)63);
  int pos = b, group_begin[5], group_tiles[4];
  auto kernel_of = [&](int i) {
    const bool aligned = (((uintptr_t)st[i].px | st[i].pitch) & 15) == 0;
    if (imgs[i].channels == 1) return 0;
    if (imgs[i].channels == 4) return 2;
    return (aligned && ctx->bulk_ok && ctx->opts.blur_mode == IRP_BLUR_EXACT) ? 3 : 1;   // the streaming kernel bakes the exact blur in
  };
  for (int g = 0; g < 4; g++) {
    group_begin[g] = pos;
    int tiles = 0;
    for (int i = b; i < e; i++) {
      if (kernel_of(i) != g) continue;
      ImgDev& d = h_imgs[pos++];
      d.px = st[i].px;
      d.pitch = st[i].pitch;
      d.w = imgs[i].width;
      d.h = imgs[i].height;
      d.c = imgs[i].channels;
      d.tiles_x = (d.w + kTileW - 1) / kTileW;
      d.tiles_y = (d.h + kTileH - 1) / kTileH;
      d.tile_base = tiles;
      d.aligned16 = (((uintptr_t)d.px | d.pitch) & 15) == 0;
      d.slot = i;
      tiles += d.tiles_x * d.tiles_y;
      if (g == 3) {
        int rc = encode_image_tmap(ctx, d, h_tmaps + (pos - 1));
        if (rc) return rc;
      }
    }
    group_tiles[g] = tiles;
  }
  group_begin[4] = pos;
  ImgDev* d_imgs = (ImgDev*)ctx->d_desc.p;
  CK(cudaMemcpyAsync(d_imgs + b, h_imgs + b, sizeof(ImgDev) * (e - b), cudaMemcpyHostToDevice, ctx->stream));
  if (group_begin[4] > group_begin[3])
    CK(cudaMemcpyAsync(d_tmaps + group_begin[3], h_tmaps + group_begin[3], sizeof(TmaDesc) * (group_begin[4] - group_begin[3]),
                       cudaMemcpyHostToDevice, ctx->stream));
  unsigned long long* d_acc = (unsigned long long*)ctx->d_acc.p;
  uint32_t* d_hist = (uint32_t*)((char*)ctx->d_acc.p + acc_bytes_for(n));
  int rc;
  if ((rc = launch_classify<1>(ctx, d_imgs + group_begin[0], group_begin[1] - group_begin[0], group_tiles[0], d_acc, d_hist))) return rc;
  if ((rc = launch_classify<3>(ctx, d_imgs + group_begin[1], group_begin[2] - group_begin[1], group_tiles[1], d_acc, d_hist))) return rc;
  if ((rc = launch_classify<4>(ctx, d_imgs + group_begin[2], group_begin[3] - group_begin[2], group_tiles[2], d_acc, d_hist))) return rc;
  if ((rc = launch_classify_bulk(ctx, d_imgs + group_begin[3], d_tmaps + group_begin[3], group_begin[4] - group_begin[3], group_tiles[3],
                                 d_acc, d_hist)))
    return rc;
  return IRP_OK;
}

// after the stream has been synchronised: integer moments -> irp_result
void finish_classify(irp_ctx* ctx, const irp_image_desc* imgs, int n, irp_result* results) {
  const unsigned long long* acc = (const unsigned long long*)ctx->h_acc.p;
  const uint32_t* hist = (const uint32_t*)((const char*)ctx->h_acc.p + acc_bytes_for(n));
  for (int i = 0; i < n; i++) {
    irp_result& r = results[i];
    memset(&r, 0, sizeof r);
    const unsigned long long* a = acc + (size_t)i * ACC_COUNT;
    for (int ch = 0; ch < 4; ch++) {
      r.sum[ch] = a[ACC_SUM + ch];
      r.sumsq[ch] = a[ACC_SUMSQ + ch];
    }
    r.e_sum[0] = a[ACC_E1S];
    r.e_sumsq[0] = a[ACC_E1Q];
    r.e_sum[1] = a[ACC_E2S];
    r.e_sumsq[1] = a[ACC_E2Q];
    r.b_sum = a[ACC_BS];
    r.b_sumsq = a[ACC_BQ];
    r.scratch_v = (uint32_t)a[ACC_SV];
    r.scratch_h = (uint32_t)a[ACC_SH];
    r.block_edges[0] = (uint32_t)a[ACC_BE0];
    r.block_edges[1] = (uint32_t)a[ACC_BE1];
    memcpy(r.luma_hist, hist + (size_t)i * 256, sizeof r.luma_hist);
    irp_scores_from_moments(&r, imgs[i].width, imgs[i].height, imgs[i].channels, imgs[i].is_jpeg);
    r.status = IRP_OK;
  }
}

// output tile: toh = 32 rows; tow = the widest of 64/32/16/8 whose source window (plus the 16-pixel
// alignment slack of the vector-load path) fits the kSrcCols columns a shared tile holds
void choose_tile(double f, int nv, int nh, int C, int* tow, int* toh, int* pairrows_max) {
  const int slack = C == 3 ? 15 : 0;
  int tw = kMaxTow;
  while (tw > 8 && (int)std::ceil(tw * f) + nh + 2 + slack > kSrcCols) tw >>= 1;
  *tow = tw;
  *toh = kMaxToh;
  int rows = (int)std::ceil(kMaxToh * f) + nv + 3;
  *pairrows_max = rows / 2 + 2;
}

template <int C>
int launch_resize(irp_ctx* ctx, const ResizeJob* d_jobs, const ResizeJob* h_jobs, int n, int total_tiles) {
  if (!n) return IRP_OK;
  size_t smem = 0;
  for (int i = 0; i < n; i++)
    smem = std::max(smem, (size_t)C * h_jobs[i].pairrows_max * kPairPitch + (size_t)C * kMaxToh * kMidPitch + 32);
  smem = round_up(smem, 16);
  if (smem > ctx->smem_optin) return fail(ctx, IRP_ERR_CUDA, "resize tile needs %zu bytes of shared memory", smem);
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, resize_kernel<C>, kResizeThreads, smem));
  if (occ < 1) return fail(ctx, IRP_ERR_CUDA, "resize kernel does not fit on an SM (%zu bytes smem)", smem);
  int grid = std::min(total_tiles, ctx->sm_count * occ);
  resize_kernel<C><<<grid, kResizeThreads, smem, ctx->stream>>>(d_jobs, n, total_tiles);
  CK(cudaGetLastError());
  ctx->timing.kernel_launches++;
  return IRP_OK;
}

template <int C>
int launch_orient(irp_ctx* ctx, const uint8_t* src, size_t spitch, int w, int h, int o, uint8_t* dst, size_t dpitch, int ow,
                  int oh) {
  if (C == 3 && ((((uintptr_t)src | spitch | (uintptr_t)dst | dpitch) & 3) == 0) && dpitch >= round_up((size_t)ow * 3, 4)) {
    dim3 g((ow + kOrTw - 1) / kOrTw, (oh + kOrTh - 1) / kOrTh);
    orient3_kernel<<<g, kOrThreads, 0, ctx->stream>>>(src, spitch, w, h, o, dst, dpitch, ow, oh);
  } else {
    dim3 b(32, 8), g((ow + 31) / 32, (oh + 7) / 8);
    orient_kernel<C><<<g, b, 0, ctx->stream>>>(src, spitch, w, h, o, dst, dpitch, ow, oh);
  }
  CK(cudaGetLastError());
  ctx->timing.kernel_launches++;
  return IRP_OK;
}

// geometry, capacity checks and staging slots for every output. mode 0: preprocess, mode 1: fusion canvas
int plan_outputs(irp_ctx* ctx, const irp_image_desc* imgs, int n, irp_out_desc* outs, int mode, std::vector<Geo>* geo,
                 std::vector<OutPlan>* oplans) {
  oplans->assign(n, OutPlan{nullptr, 0, false});
  geo->assign(n, Geo{});
  size_t orient_total = 0, out_total = 0;
  for (int i = 0; i < n; i++) {
    if (!imgs[i].pixels) continue;  // unused fusion slot
    const irp_image_desc& d = imgs[i];
    Geo& g = (*geo)[i];
    g.o = (d.exif_orientation >= 1 && d.exif_orientation <= 8) ? d.exif_orientation : 1;
    orient_dims(d.width, d.height, g.o, &g.wo, &g.ho);
    g.ox = g.oy = 0;
    if (mode == 0) {
      preprocess_dims(d.width, d.height, g.o, &g.dw, &g.dh, &g.f);
      g.dc = d.channels == 1 ? 1 : 3;
    } else {
      fusion_dims(d.width, d.height, g.o, &g.dw, &g.dh, &g.ox, &g.oy, &g.f);
      g.dc = 3;
    }
    g.kh = std::max(1, (int)std::floor((double)g.wo / g.dw / 2.0));
    g.kv = std::max(1, (int)std::floor((double)g.ho / g.dh / 2.0));
    g.rw = (g.wo + g.kh - 1) / g.kh;
    g.rh = (g.ho + g.kv - 1) / g.kv;
    g.fh = g.f / g.kh;
    g.fv = g.f / g.kv;
    int out_w = mode == 0 ? g.dw : IRP_FUSION_CANVAS, out_h = mode == 0 ? g.dh : IRP_FUSION_CANVAS;
    irp_out_desc& od = outs[i];
    if (!od.pixels) return fail(ctx, IRP_ERR_BAD_ARG, "output %d: null pixels", i);
    size_t tight = (size_t)out_w * g.dc;
    size_t pitch = od.pitch ? od.pitch : tight;
    if (pitch < tight) return fail(ctx, IRP_ERR_BAD_ARG, "output %d: pitch %zu < %zu", i, pitch, tight);
    if (od.capacity < pitch * (size_t)(out_h - 1) + tight)
      return fail(ctx, IRP_ERR_CAPACITY, "output %d: capacity %zu too small for %dx%dx%d", i, od.capacity, out_w, out_h, g.dc);
    od.width = out_w;
    od.height = out_h;
    od.channels = g.dc;
    if (g.o != 1) {
      g.orient_off = orient_total;
      orient_total += round_up(round_up((size_t)g.wo * d.channels, 16) * g.ho, 256);
    }
    if (g.kh > 1 || g.kv > 1) {
      g.box_off = orient_total;
      orient_total += round_up(round_up((size_t)g.rw * d.channels, 16) * g.rh, 256);
    }
    if (!od.on_device) {   // staged rows get a 16-byte pitch: the tensor-core resize stores whole 8-byte pieces
      g.out_off = out_total;
      out_total += round_up(round_up(tight, 16) * out_h, 256);
    }
  }
  if (orient_total) CK(ctx->d_orient.reserve(orient_total));
  if (out_total) CK(ctx->d_stage_out.reserve(out_total));
  for (int i = 0; i < n; i++) {
    if (!imgs[i].pixels) continue;
    const Geo& g = (*geo)[i];
    const irp_out_desc& od = outs[i];
    OutPlan& op = (*oplans)[i];
    if (od.on_device) {
      op = OutPlan{od.pixels, od.pitch ? od.pitch : (size_t)od.width * g.dc, false};
    } else {
      op = OutPlan{(uint8_t*)ctx->d_stage_out.p + g.out_off, round_up((size_t)od.width * g.dc, 16), true};
    }
  }
  return IRP_OK;
}

// the streaming kernel (irp_resize_tma.cuh): smallest tile footprint that serves every tile of a job
struct RtFoot { int tow, toh, ncols, nrows; };
RtLayout rt_layout(int ncols_px, int nrows, int groups, int mid_rows);
// Largest output tile whose source footprint fits the TMA box limits (512 bytes x 256 rows) and, together
// with the boxes already chosen for this launch, one group's share of shared memory.
bool rt_footprint(const PlanDev& v, const PlanDev& h, int dw, int dh, int launch_cols, int launch_rows, size_t budget, int groups, int toh_max,
                  RtFoot* f) {
  const int cand[][2] = {{64, toh_max}, {64, 16}, {32, toh_max}, {32, 16}, {64, 8}, {16, toh_max}, {32, 8}, {16, 16}, {16, 8}};
  for (const auto& c : cand) {
    const int tow = c[0], toh = c[1];
    int ncols = 0, nrows = 0;
    for (int ox0 = 0; ox0 < dw; ox0 += tow) {
      const int last = std::min(ox0 + tow, dw) - 1;
      ncols = std::max(ncols, h.h_start[last] + h.n - (h.h_start[ox0] & ~3));
    }
    for (int oy0 = 0; oy0 < dh; oy0 += toh) {
      const int last = std::min(oy0 + toh, dh) - 1;
      nrows = std::max(nrows, v.h_start[last] + v.n - (v.h_start[oy0] & ~1));
    }
    const int ntr = (ncols + 3) / 4, rows = (nrows + 1) & ~1;
    if (ntr * 12 + 12 > 512 || rows > 256) continue;
    const RtLayout L = rt_layout(std::max(launch_cols, ntr * 4), std::max(launch_rows, rows), groups, toh_max);
    if ((size_t)L.group_bytes > budget || L.box_cols > 512 || L.box_rows > 256) continue;
    *f = RtFoot{tow, toh, ntr * 4, rows};
    return true;
  }
  return false;
}

RtLayout rt_layout(int ncols_px, int nrows, int groups, int mid_rows) {
  RtLayout L;
  L.groups = groups;
  L.mid_rows = mid_rows;
  L.box_cols = (int)round_up((size_t)ncols_px * 3 + 12, 16);   // + the sub-16-byte offset of the first pixel
  L.box_rows = nrows;
  L.mid_pitch = (int)round_up((size_t)ncols_px + 8, 16);
  size_t p = round_up((size_t)L.box_cols * L.box_rows, 128);
  L.off_mid = (int)p;
  p += (size_t)3 * mid_rows * L.mid_pitch;
  p = round_up(p, 128);
  L.off_vtab = (int)p;
  p += (size_t)mid_rows * kTabWords * 4;
  L.off_hcols = (int)p;
  p += (size_t)kRTow * kHTabWords * 4;
  L.off_sync = (int)p;
  p += 128;
  L.group_bytes = (int)round_up(p, 128);
  return L;
}

int encode_source_tmap(irp_ctx* ctx, const uint8_t* px, size_t pitch, int w, int h, int box_cols, int box_rows, TmaDesc* out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                               CUtensorMapFloatOOBfill);
  const cuuint64_t dims[2] = {(cuuint64_t)(w * 3 + 1) / 2, (cuuint64_t)h};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {(cuuint32_t)box_cols / 2, (cuuint32_t)box_rows}, estr[2] = {1, 1};
  CUresult r = ((EncodeFn)ctx->encode_tiled)(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, (void*)px, dims, strides,
                                             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, IRP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %dx%d source, pitch %zu", (int)r, w, h, pitch);
  return IRP_OK;
}

// shared-memory map of the tensor-core resize kernel.  The source buffers come first: the last K step of a quarter may
// read rows past its buffer (times zero coefficients), and what lies behind has to be mapped memory.
MmLayout mm_layout(int rows, int ksv, int nbuf) {
  MmLayout L;
  L.R = rows;
  L.ksv_max = ksv;
  L.nbuf = nbuf;
  size_t p = 0;
  L.off_src = (int)p;
  p += (size_t)nbuf * 2 * rows * 128;
  L.off_cv = (int)p;
  p += (size_t)4 * ksv * 2048;
  L.off_mid = (int)p;
  p += (size_t)3 * kMmMidPlane;
  L.off_ch = (int)p;
  p += (size_t)2 * kMmChBytes;
  L.off_out = (int)p;
  p += (size_t)kMmTR * kMmOutPitch;
  L.off_info = (int)p;
  p += 2 * sizeof(MmInfo) + 32;
  p = round_up(p, 16);
  L.off_bar = (int)p;
  p += 8 * kBarCount;
  L.total = (int)p;
  return L;
}
int encode_mm_tmap(irp_ctx* ctx, const uint8_t* px, size_t pitch, int w, int h, int box_rows, TmaDesc* out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                               CUtensorMapFloatOOBfill);
  const cuuint64_t dims[2] = {(cuuint64_t)w * 3, (cuuint64_t)h};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch};
  cuuint32_t box[2] = {128, (cuuint32_t)box_rows}, estr[2] = {1, 1};
  CUresult r = ((EncodeFn)ctx->encode_tiled)(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)px, dims, strides,
                                             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, IRP_ERR_CUDA, "cuTensorMapEncodeTiled (swizzled) failed (%d) for a %dx%d source, pitch %zu", (int)r, w, h, pitch);
  return IRP_OK;
}

// orient + resize images [b, e); job slots [b, e) of the batch-wide arrays.  Jobs are grouped by
// kernel: 1 / 3 / 4 channels on the generic kernel, and the streaming kernel for 3-channel sources
// with aligned rows and a real shrink.
int resize_range(irp_ctx* ctx, const irp_image_desc* imgs, const std::vector<Staged>& st, const std::vector<Geo>& geo,
                 const std::vector<OutPlan>& oplans, irp_out_desc* outs, int mode, int b, int e) {
  ResizeJob* h_jobs = (ResizeJob*)ctx->h_jobs.p;
  RtJob* h_rt = (RtJob*)ctx->h_rtjobs.p;
  TmaDesc* h_tm = (TmaDesc*)(((uintptr_t)ctx->h_rtmaps.p + 63) & ~(uintptr_t)63);
  TmaDesc* d_tm = (TmaDesc*)(((uintptr_t)ctx->d_rtmaps.p + 63) & ~(uintptr_t)63);
  struct Src { const uint8_t* px; size_t pitch; };
  std::vector<Src> src(e - b);
  std::vector<int> kern(e - b, -1);
  {
    // Geometry plans (tap tables, the tensor-core kernel's operand matrices: up to ~1 MB per axis) are cached for the
    // life of the context; uploads of ever-new sizes must not grow that without bound.  Past the cap the whole cache
    // goes (nothing of it is in flight: plans are only referenced by this function's launches, drained here).
    static const size_t cap = (size_t)(getenv("IRP_PLAN_CACHE_MB") ? std::max(1, atoi(getenv("IRP_PLAN_CACHE_MB"))) : 2048) << 20;
    if (ctx->plan_total > cap) {
      CK(wait_stream(ctx, ctx->stream));
      for (void* p : ctx->plan_chunks) cudaFree(p);
      ctx->plan_chunks.clear();
      ctx->plans.clear();
      ctx->plan_total = ctx->plan_used = ctx->plan_cap = 0;
    }
  }
  std::vector<RtFoot> foot(e - b);
  std::vector<AxisPlan> pv(e - b), ph(e - b);
  std::vector<const PlanDev*> dv(e - b), dh(e - b);
  int rc;
  // pass 1: orientation, plans
  for (int i = b; i < e; i++) {
    if (!imgs[i].pixels) continue;
    const irp_image_desc& d = imgs[i];
    const Geo& g = geo[i];
    Src& s = src[i - b];
    if (g.o != 1) {
      uint8_t* op = (uint8_t*)ctx->d_orient.p + g.orient_off;
      size_t opitch = round_up((size_t)g.wo * d.channels, 16);
      switch (d.channels) {
        case 1: rc = launch_orient<1>(ctx, st[i].px, st[i].pitch, d.width, d.height, g.o, op, opitch, g.wo, g.ho); break;
        case 3: rc = launch_orient<3>(ctx, st[i].px, st[i].pitch, d.width, d.height, g.o, op, opitch, g.wo, g.ho); break;
        default: rc = launch_orient<4>(ctx, st[i].px, st[i].pitch, d.width, d.height, g.o, op, opitch, g.wo, g.ho); break;
      }
      if (rc) return rc;
      s = Src{op, opitch};
    } else {
      s = Src{st[i].px, st[i].pitch};
    }
    if (g.kh > 1 || g.kv > 1) {   // integer box pre-shrink; the lanczos passes then do the remaining factor in [2, 4)
      uint8_t* bp = (uint8_t*)ctx->d_orient.p + g.box_off;
      const size_t bpitch = round_up((size_t)g.rw * d.channels, 16);
      const dim3 grid((g.rw * d.channels + 255) / 256, g.rh);
      switch (d.channels) {
        case 1: box_shrink_kernel<1><<<grid, 256, 0, ctx->stream>>>(s.px, s.pitch, g.wo, g.ho, g.kh, g.kv, bp, bpitch, g.rw, g.rh); break;
        case 3: box_shrink_kernel<3><<<grid, 256, 0, ctx->stream>>>(s.px, s.pitch, g.wo, g.ho, g.kh, g.kv, bp, bpitch, g.rw, g.rh); break;
        default: box_shrink_kernel<4><<<grid, 256, 0, ctx->stream>>>(s.px, s.pitch, g.wo, g.ho, g.kh, g.kv, bp, bpitch, g.rw, g.rh); break;
      }
      CK(cudaGetLastError());
      ctx->timing.kernel_launches++;
      s = Src{bp, bpitch};
    }
    if ((rc = get_plan(ctx, g.rh, g.dh, g.fv, &pv[i - b], &dv[i - b]))) return rc;
    if ((rc = get_plan(ctx, g.rw, g.dw, g.fh, &ph[i - b], &dh[i - b]))) return rc;
  }
  // kernel choice.  The streaming kernel runs 5 groups x 24-row tiles per CTA when every eligible job keeps
  // a full-size tile inside a fifth of the shared memory, else 4 groups x 32-row tiles.
  int box_cols_px = 0, box_rows = 0, rt_groups = kRMaxGroups, rt_toh = 24;
  int mm_fit_rows[2], mm_fit_ksv[2];
  for (int attempt = 0; attempt < 2; attempt++) {
    box_cols_px = box_rows = 0;
    mm_fit_rows[0] = mm_fit_rows[1] = 16;
    mm_fit_ksv[0] = mm_fit_ksv[1] = 1;
    const size_t budget = ((size_t)ctx->smem_optin_full - 4096) / rt_groups;
    bool degraded = false;
    for (int i = b; i < e; i++) {
      if (!imgs[i].pixels) continue;
      const irp_image_desc& d = imgs[i];
      const Geo& g = geo[i];
      const Src& s = src[i - b];
      int k = d.channels == 1 ? 0 : (d.channels == 4 ? 2 : 1);
      if (k == 1 && ctx->bulk_ok && ctx->rtma_ok && ((((uintptr_t)s.px | s.pitch) & 15) == 0) && pv[i - b].n > 1 && ph[i - b].n > 1) {
        if (rt_footprint(*dv[i - b], *dh[i - b], g.dw, g.dh, box_cols_px, box_rows, budget, rt_groups, rt_toh, &foot[i - b])) {
          k = 3;
          box_cols_px = std::max(box_cols_px, foot[i - b].ncols);
          box_rows = std::max(box_rows, foot[i - b].nrows);
          degraded |= foot[i - b].tow < kRTow || foot[i - b].toh < rt_toh;
        } else {
          degraded = true;
        }
      }
      // no geometry change and nothing to normalise: the oriented pixels ARE the result, a row copy
      if (pv[i - b].n == 1 && ph[i - b].n == 1 && g.dc == d.channels && g.dw == g.rw && g.dh == g.rh) k = 4;
      // the tensor-core kernel: 3-channel aligned sources whose geometry fits its tile, 8-byte aligned destination pieces
      if ((k == 1 || k == 3) && ctx->rmma_ok && ((((uintptr_t)s.px | s.pitch) & 15) == 0) && dv[i - b]->mm_v_ok && dh[i - b]->mm_h_ok &&
          ((((uintptr_t)oplans[i].dev | oplans[i].dev_pitch) & 15) == 0) && ((3 * g.ox) & 7) == 0) {
        // a launch's layout is sized for the largest rows and K steps among its jobs: a job joins the double-buffered
        // launch if the combined layout still fits, else the single-buffered one, else it stays with the ALU kernels
        for (int mg = 0; mg < 2; mg++) {
          const int r2 = std::max(mm_fit_rows[mg], dv[i - b]->mm_rows), k2 = std::max(mm_fit_ksv[mg], dv[i - b]->mm_ksv);
          if (mm_layout(r2, k2, mg == 0 ? 2 : 1).total <= (int)ctx->smem_optin_full - 2048) {
            mm_fit_rows[mg] = r2;
            mm_fit_ksv[mg] = k2;
            k = 5 + mg;
            break;
          }
        }
      }
      kern[i - b] = k;
    }
    if (!degraded || attempt == 1) break;
    rt_groups = 4;
    rt_toh = kRTohMax;
  }
  const RtLayout L = rt_layout(std::max(box_cols_px, 4), std::max(box_rows, 2), rt_groups, rt_toh);
  if (getenv("IRP_TRACE")) fprintf(stderr, "resize_tma: groups %d toh %d box %d x %d group_bytes %d\n", L.groups, rt_toh, L.box_cols, L.box_rows, L.group_bytes);
  // pass 2: job descriptors, grouped by kernel
  int pos = b, group_begin[8], group_tiles[7];
  int mm_rows[2] = {16, 16}, mm_ksv[2] = {1, 1};
  for (int gi = 0; gi < 7; gi++) {
    group_begin[gi] = pos;
    int tiles = 0;
    for (int i = b; i < e; i++) {
      if (!imgs[i].pixels || kern[i - b] != gi) continue;
      const irp_image_desc& d = imgs[i];
      const Geo& g = geo[i];
      const irp_out_desc& od = outs[i];
      const OutPlan& op = oplans[i];
      if (mode == 1)  // black pad: clear the whole canvas, the tiles then write the image once
        CK(cudaMemset2DAsync(op.dev, op.dev_pitch, 0, (size_t)od.width * g.dc, od.height, ctx->stream));
      if (gi == 4) {   // identity geometry: a row copy (into the canvas window for fusion); descriptors share the job buffer
        CopyJob& K = reinterpret_cast<CopyJob*>(h_jobs + pos)[0];
        static_assert(sizeof(CopyJob) <= sizeof(ResizeJob), "copy descriptors are stored in resize-job slots");
        K.src = src[i - b].px;
        K.spitch = src[i - b].pitch;
        K.dst = op.dev + (size_t)g.oy * op.dev_pitch + (size_t)g.ox * g.dc;
        K.dpitch = op.dev_pitch;
        K.row_bytes = g.dw * g.dc;
        K.rows = g.dh;
        K.row_base = tiles;
        K.pad = 0;
        tiles += g.dh;
        pos++;
        continue;
      }
      if (gi >= 5) {
        MmJob& M = ((MmJob*)ctx->h_mmjobs.p)[pos];
        memset(&M, 0, sizeof M);
        M.dst = op.dev;
        M.dst_pitch = op.dev_pitch;
        M.vfirst = dv[i - b]->mm_first; M.vmat = dv[i - b]->mm_vmat; M.vmats = dv[i - b]->mm_vmats;
        M.hfirst = dh[i - b]->mm_first; M.hmat = dh[i - b]->mm_hmat; M.hmats = dh[i - b]->mm_hmats;
        M.sw = g.rw; M.sh = g.rh; M.dw = g.dw; M.dh = g.dh;
        M.dst_x0 = g.ox; M.dst_y0 = g.oy;
        M.tiles_x = (g.dw + kMmTC - 1) / kMmTC;
        M.tiles_y = (g.dh + kMmTR - 1) / kMmTR;
        M.strip_base = tiles;   // this group counts strips
        M.ksv = dv[i - b]->mm_ksv;
        tiles += M.tiles_x;
        mm_rows[gi - 5] = std::max(mm_rows[gi - 5], dv[i - b]->mm_rows);
        mm_ksv[gi - 5] = std::max(mm_ksv[gi - 5], dv[i - b]->mm_ksv);
        pos++;
        continue;
      }
      if (gi == 3) {
        RtJob& R = h_rt[pos];
        memset(&R, 0, sizeof R);
        R.dst = op.dev;
        R.dst_pitch = op.dev_pitch;
        R.vrows = dv[i - b]->vrows;
        R.hcols = dh[i - b]->hcols;
        R.vstart = dv[i - b]->start;
        R.hstart = dh[i - b]->start;
        R.sw = g.rw; R.sh = g.rh; R.dw = g.dw; R.dh = g.dh;
        R.dst_x0 = g.ox; R.dst_y0 = g.oy;
        R.tow = foot[i - b].tow; R.toh = foot[i - b].toh;
        R.tiles_x = (g.dw + R.tow - 1) / R.tow;
        R.tiles_y = (g.dh + R.toh - 1) / R.toh;
        R.tile_base = tiles;
        R.vn = pv[i - b].n; R.hn = ph[i - b].n;
        tiles += R.tiles_x * R.tiles_y;
        R.box_cols = (int)round_up((size_t)foot[i - b].ncols * 3 + 12, 16);
        R.box_rows = foot[i - b].nrows;
        if ((rc = encode_source_tmap(ctx, src[i - b].px, src[i - b].pitch, g.rw, g.rh, R.box_cols, R.box_rows, h_tm + pos))) return rc;
        pos++;
        continue;
      }
      ResizeJob& J = h_jobs[pos++];
      memset(&J, 0, sizeof J);
      J.src = src[i - b].px;
      J.src_pitch = src[i - b].pitch;
      J.dst = op.dev;
      J.dst_pitch = op.dev_pitch;
      J.sw = g.rw;
      J.sh = g.rh;
      J.c = d.channels;
      J.dc = g.dc;
      J.dw = g.dw;
      J.dh = g.dh;
      J.dst_x0 = g.ox;
      J.dst_y0 = g.oy;
      J.expand_grey = (d.channels == 1 && g.dc == 3);
      J.aligned16 = (((uintptr_t)J.src | J.src_pitch) & 15) == 0;
      J.v = pv[i - b];
      J.h = ph[i - b];
      choose_tile(std::max(g.fh, g.fv), J.v.n, J.h.n, d.channels, &J.tow, &J.toh, &J.pairrows_max);
      J.tiles_x = (J.dw + J.tow - 1) / J.tow;
      J.tiles_y = (J.dh + J.toh - 1) / J.toh;
      J.tile_base = tiles;
      tiles += J.tiles_x * J.tiles_y;
    }
    group_tiles[gi] = tiles;
  }
  group_begin[7] = pos;
  if (pos == b) return IRP_OK;
  ResizeJob* d_jobs = (ResizeJob*)ctx->d_jobs.p;
  if (group_begin[3] > b)
    CK(cudaMemcpyAsync(d_jobs + b, h_jobs + b, sizeof(ResizeJob) * (group_begin[3] - b), cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = launch_resize<1>(ctx, d_jobs + group_begin[0], h_jobs + group_begin[0], group_begin[1] - group_begin[0], group_tiles[0]))) return rc;
  if ((rc = launch_resize<3>(ctx, d_jobs + group_begin[1], h_jobs + group_begin[1], group_begin[2] - group_begin[1], group_tiles[1]))) return rc;
  if ((rc = launch_resize<4>(ctx, d_jobs + group_begin[2], h_jobs + group_begin[2], group_begin[3] - group_begin[2], group_tiles[2]))) return rc;
  if (const int ncp = group_begin[5] - group_begin[4]) {
    // copy jobs sit in ResizeJob-sized slots: compact them into a dense CopyJob array at the front of their range
    const int g4 = group_begin[4];
    CopyJob* hc = reinterpret_cast<CopyJob*>(h_jobs + g4);
    for (int k = 1; k < ncp; k++) hc[k] = reinterpret_cast<CopyJob*>(h_jobs + g4 + k)[0];
    CopyJob* dc = reinterpret_cast<CopyJob*>(d_jobs + g4);
    CK(cudaMemcpyAsync(dc, hc, sizeof(CopyJob) * ncp, cudaMemcpyHostToDevice, ctx->stream));
    const int rows = group_tiles[4];
    const int grid = std::min((rows + 7) / 8, ctx->sm_count * 8);
    copy_rows_kernel<<<grid, 256, 0, ctx->stream>>>(dc, ncp, rows);
    CK(cudaGetLastError());
    ctx->timing.kernel_launches++;
  }
  if (const int nrt = group_begin[4] - group_begin[3]) {
    RtJob* d_rt = (RtJob*)ctx->d_rtjobs.p;
    const int g3 = group_begin[3];
    CK(cudaMemcpyAsync(d_rt + g3, h_rt + g3, sizeof(RtJob) * nrt, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_tm + g3, h_tm + g3, sizeof(TmaDesc) * nrt, cudaMemcpyHostToDevice, ctx->stream));
    const size_t smem = (size_t)L.group_bytes * L.groups + 128;
    const int grid = std::min((group_tiles[3] + L.groups - 1) / L.groups, ctx->sm_count);
    resize_tma_kernel<<<grid, L.groups * 128, smem, ctx->stream>>>(d_rt + g3, d_tm + g3, nrt, group_tiles[3], L);
    CK(cudaGetLastError());
    ctx->timing.kernel_launches++;
  }
  for (int mg = 0; mg < 2; mg++) {   // the tensor-core kernel: jobs whose footprint fits twice, then those that fit once
    const int g5 = group_begin[5 + mg], nmm = group_begin[6 + mg] - g5, strips = group_tiles[5 + mg];
    if (!nmm) continue;
    // every job's source as a tensor map of 128-byte x (rows / 2) boxes with the 128-byte swizzle
    const MmLayout ML = mm_layout(mm_rows[mg], mm_ksv[mg], mg == 0 ? 2 : 1);
    MmJob* h_mm = (MmJob*)ctx->h_mmjobs.p;
    MmJob* d_mm = (MmJob*)ctx->d_mmjobs.p;
    TmaDesc* h_mt = (TmaDesc*)(((uintptr_t)ctx->h_mmmaps.p + 63) & ~(uintptr_t)63);
    TmaDesc* d_mt = (TmaDesc*)(((uintptr_t)ctx->d_mmmaps.p + 63) & ~(uintptr_t)63);
    int slot = g5;
    for (int i = b; i < e; i++) {
      if (!imgs[i].pixels || kern[i - b] != 5 + mg) continue;
      if ((rc = encode_mm_tmap(ctx, src[i - b].px, src[i - b].pitch, geo[i].rw, geo[i].rh, ML.R / 2, h_mt + slot))) return rc;
      slot++;
    }
    CK(cudaMemcpyAsync(d_mm + g5, h_mm + g5, sizeof(MmJob) * nmm, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_mt + g5, h_mt + g5, sizeof(TmaDesc) * nmm, cudaMemcpyHostToDevice, ctx->stream));
    const int grid = std::min(strips, ctx->sm_count);
    if (getenv("IRP_TRACE"))
      fprintf(stderr, "resize_mma: %d jobs, %d strips, rows %d, ksv %d, %d source buffer(s), smem %d\n", nmm, strips, ML.R, ML.ksv_max, ML.nbuf, ML.total);
    if (auto e2 = ctx->d_mmctr.reserve(256); e2 != cudaSuccess) return fail(ctx, IRP_ERR_NOMEM, "strip counter: %s", cudaGetErrorString(e2));
    int* counter = (int*)ctx->d_mmctr.p + 16 * mg;
    CK(cudaMemsetAsync(counter, 0, 4, ctx->stream));
    long long* dbg = nullptr;
    if (getenv("IRP_MMA_DEBUG")) {
      CK(cudaMalloc(&dbg, 80 * 8));
      CK(cudaMemsetAsync(dbg, 0, 80 * 8, ctx->stream));
    }
    resize_mma_kernel<<<grid, kMmThreads, ML.total + 1024, ctx->stream>>>(d_mm + g5, d_mt + g5, nmm, strips, counter, ML, dbg);
    CK(cudaGetLastError());
    ctx->timing.kernel_launches++;
    if (dbg) {   // cycles block 0 spent waiting, per role and barrier kind (kernel experiments only)
      long long h[80];
      CK(wait_stream(ctx, ctx->stream));
      CK(cudaMemcpy(h, dbg, sizeof h, cudaMemcpyDeviceToHost));
      cudaFree(dbg);
      const char* role[5] = {"epilogue w0 [VFull MidFree HFull Full - - | EV: ld st+pack arrive sts | EH: tmem global]", "producer [SrcFree CvFree ChFree]",
                             "issuer 0 [VFree CvFull MidFull HFree ChFull Full]", "issuer 1", "issuer 2"};
      for (int r = 0; r < 5; r++) {
        fprintf(stderr, "resize_mma waits, %s:", role[r]);
        for (int k = 0; k < 12; k++) fprintf(stderr, " %lld", h[16 * r + k]);
        fprintf(stderr, "  of %lld cycles\n", h[16 * r + 15]);
      }
    }
  }
  return IRP_OK;
}

int copy_outputs(irp_ctx* ctx, const irp_image_desc* imgs, irp_out_desc* outs, const std::vector<OutPlan>& oplans, int b, int e,
                 cudaStream_t stream) {
  for (int i = b; i < e; i++) {
    if (!imgs[i].pixels || !oplans[i].via_stage) continue;
    irp_out_desc& od = outs[i];
    size_t tight = (size_t)od.width * od.channels;
    CK(copy_rows_async(od.pixels, od.pitch ? od.pitch : tight, oplans[i].dev, oplans[i].dev_pitch, tight, od.height,
                       cudaMemcpyDeviceToHost, stream));
  }
  return IRP_OK;
}

// The one driver behind classify / preprocess / analyze / fusion.
// Host-resident batches are cut into chunks (~256 MB of pixels) and pipelined over three streams:
// H2D of chunk c+1, the kernels of chunk c and the D2H of chunk c-1 overlap, so the call is bound by
// the slower PCIe direction rather than by the sum of copies and kernels.  Device-resident batches
// are one chunk: one launch per kernel and channel count.
int run_batch_locked(irp_ctx* ctx, const irp_image_desc* imgs, int n, irp_result* results, irp_out_desc* outs, int resize_mode);
int run_batch(irp_ctx* ctx, const irp_image_desc* imgs, int n, irp_result* results, irp_out_desc* outs, int resize_mode) {
  if (!ctx) return IRP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->err.clear();
  return run_batch_locked(ctx, imgs, n, results, outs, resize_mode);
}
// A failing call must not leave copies to or from the caller's buffers (or reads of the context's pinned
// scratch) in flight: the caller may free or reuse them as soon as the call returns.
void drain_streams(irp_ctx* ctx) {
  cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_in_stream) cudaStreamSynchronize(ctx->copy_in_stream);
  if (ctx->copy_out_stream) cudaStreamSynchronize(ctx->copy_out_stream);
}
int run_batch_inner(irp_ctx* ctx, const irp_image_desc* imgs, int n, irp_result* results, irp_out_desc* outs, int resize_mode);
int run_batch_locked(irp_ctx* ctx, const irp_image_desc* imgs, int n, irp_result* results, irp_out_desc* outs, int resize_mode) {
  const int rc = run_batch_inner(ctx, imgs, n, results, outs, resize_mode);
  if (rc != IRP_OK) drain_streams(ctx);
  return rc;
}
int run_batch_inner(irp_ctx* ctx, const irp_image_desc* imgs, int n, irp_result* results, irp_out_desc* outs, int resize_mode) {
  if (n < 0 || (n > 0 && !imgs)) return fail(ctx, IRP_ERR_BAD_ARG, "bad batch arguments");
  if (n == 0) return IRP_OK;
  CK(cudaSetDevice(ctx->device));
  int rc;
  for (int i = 0; i < n; i++) {
    if (resize_mode == 1 && !imgs[i].pixels) continue;
    if ((rc = validate_desc(ctx, imgs[i], i))) return rc;
  }
  ctx->timing = irp_timing{};
  std::vector<Staged> st;
  std::vector<Geo> geo;
  std::vector<OutPlan> oplans;
  if ((rc = plan_inputs(ctx, imgs, n, &st))) return rc;
  if (outs && (rc = plan_outputs(ctx, imgs, n, outs, resize_mode, &geo, &oplans))) return rc;
  if (results) {
    CK(ctx->d_desc.reserve(sizeof(ImgDev) * n));
    CK(ctx->h_desc.reserve(sizeof(ImgDev) * n));
    CK(ctx->d_tmaps.reserve(sizeof(TmaDesc) * n + 64));
    CK(ctx->h_tmaps.reserve(sizeof(TmaDesc) * n + 64));
    CK(ctx->d_acc.reserve(acc_bytes_for(n) + sizeof(uint32_t) * 256 * n));
    CK(ctx->h_acc.reserve(acc_bytes_for(n) + sizeof(uint32_t) * 256 * n));
  }
  if (outs) {
    CK(ctx->d_jobs.reserve(sizeof(ResizeJob) * n));
    CK(ctx->h_jobs.reserve(sizeof(ResizeJob) * n));
    CK(ctx->d_rtjobs.reserve(sizeof(RtJob) * n));
    CK(ctx->h_rtjobs.reserve(sizeof(RtJob) * n));
    CK(ctx->d_rtmaps.reserve(sizeof(TmaDesc) * n + 64));
    CK(ctx->h_rtmaps.reserve(sizeof(TmaDesc) * n + 64));
    CK(ctx->d_mmjobs.reserve(sizeof(MmJob) * n));
    CK(ctx->h_mmjobs.reserve(sizeof(MmJob) * n));
    CK(ctx->d_mmmaps.reserve(sizeof(TmaDesc) * n + 64));
    CK(ctx->h_mmmaps.reserve(sizeof(TmaDesc) * n + 64));
  }
  // chunk boundaries
  std::vector<int> cuts{0};
  {
    bool any_host = false;
    for (int i = 0; i < n; i++) any_host |= st[i].bytes > 0 || (outs && imgs[i].pixels && oplans[i].via_stage);
    if (any_host) {
      size_t acc = 0;
      for (int i = 0; i < n; i++) {
        acc += st[i].bytes + (outs && imgs[i].pixels && oplans[i].via_stage ? (size_t)outs[i].width * outs[i].height * outs[i].channels : 0);
        if (acc >= ctx->chunk_bytes && i + 1 < n) {
          cuts.push_back(i + 1);
          acc = 0;
        }
      }
    }
    cuts.push_back(n);
  }
  const int nchunks = (int)cuts.size() - 1;
  const bool piped = nchunks > 1;
  cudaStream_t s_in = piped ? ctx->copy_in_stream : ctx->stream, s_out = piped ? ctx->copy_out_stream : ctx->stream;

  CK(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (piped) {  // the copy streams must not start before work already queued on the caller's stream
    CK(cudaStreamWaitEvent(s_in, ctx->ev[0], 0));
    CK(cudaStreamWaitEvent(s_out, ctx->ev[0], 0));
  }
  if (results) CK(cudaMemsetAsync(ctx->d_acc.p, 0, acc_bytes_for(n) + sizeof(uint32_t) * 256 * n, ctx->stream));
  for (int c = 0; c < nchunks; c++) {
    const int b = cuts[c], e = cuts[c + 1];
    cudaEvent_t ev_in, ev_done, t0, t1, t2;
    if ((rc = copy_inputs(ctx, imgs, st, b, e, s_in))) return rc;
    if (piped) {
      CK(get_event(ctx, 2 * c, &ev_in, false));
      CK(cudaEventRecord(ev_in, s_in));
      CK(cudaStreamWaitEvent(ctx->stream, ev_in, 0));
    }
    CK(get_event(ctx, 3 * c, &t0, true));
    CK(get_event(ctx, 3 * c + 1, &t1, true));
    CK(get_event(ctx, 3 * c + 2, &t2, true));
    CK(cudaEventRecord(t0, ctx->stream));
    if (results && (rc = classify_range(ctx, imgs, st, n, b, e))) return rc;
    CK(cudaEventRecord(t1, ctx->stream));
    if (outs && (rc = resize_range(ctx, imgs, st, geo, oplans, outs, resize_mode, b, e))) return rc;
    CK(cudaEventRecord(t2, ctx->stream));
    if (piped) {
      CK(get_event(ctx, 2 * c + 1, &ev_done, false));
      CK(cudaEventRecord(ev_done, ctx->stream));
      CK(cudaStreamWaitEvent(s_out, ev_done, 0));
    }
    if (outs && (rc = copy_outputs(ctx, imgs, outs, oplans, b, e, s_out))) return rc;
  }
  if (results) {
    CK(cudaMemcpyAsync(ctx->h_acc.p, ctx->d_acc.p, acc_bytes_for(n) + sizeof(uint32_t) * 256 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_error_flag, ctx->d_error_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (piped) {  // join the output stream back into the caller's stream
    CK(cudaEventRecord(ctx->ev[5], s_out));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev[5], 0));
  }
  CK(cudaEventRecord(ctx->ev[4], ctx->stream));
  CK(wait_stream(ctx, ctx->stream));
  {
    float ms = 0, sum_c = 0, sum_p = 0;
    for (int c = 0; c < nchunks; c++) {
      cudaEventElapsedTime(&ms, ctx->timing_events[3 * c], ctx->timing_events[3 * c + 1]);
      sum_c += ms;
      cudaEventElapsedTime(&ms, ctx->timing_events[3 * c + 1], ctx->timing_events[3 * c + 2]);
      sum_p += ms;
    }
    ctx->timing.classify_ms = sum_c;
    ctx->timing.preprocess_ms = sum_p;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[4]);
    ctx->timing.total_ms = ms;
    if (!piped) {
      cudaEventElapsedTime(&ms, ctx->ev[0], ctx->timing_events[0]);
      ctx->timing.h2d_ms = ms;
      cudaEventElapsedTime(&ms, ctx->timing_events[2], ctx->ev[4]);
      ctx->timing.d2h_ms = ms;
    }
    ctx->timing.chunks = (uint32_t)nchunks;
  }
  if (results && *ctx->h_error_flag)
    return fail(ctx, IRP_ERR_CUDA, "classify kernel: shared-memory layout does not fit the probed allocation");
  if (results) finish_classify(ctx, imgs, n, results);
  return IRP_OK;
}

#include "irp_icc.inc"

// the profile an encode call attaches: id 0 = the context default (may be empty)
std::shared_ptr<const std::vector<uint8_t>> resolve_icc(irp_ctx* ctx, int id) {
  irp_ctx* root = ctx->parent ? ctx->parent : ctx;
  std::lock_guard<std::mutex> lk(root->icc_mu);
  if (id < 0 || id >= (int)root->icc_profiles.size()) return nullptr;
  return root->icc_profiles[id];
}

#include "irp_jpeg_host.inc"
#include "irp_jpeg_enc_host.inc"

}  // namespace

// ============================================================================
// C ABI
// ============================================================================
extern "C" {

int irp_abi_version(void) { return IRP_ABI_VERSION; }

int irp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

irp_ctx* irp_create(int device, const irp_opts* opts) {
  irp_ctx* ctx = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    fail(nullptr, IRP_ERR_NO_DEVICE, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    return nullptr;
  }
  if (device < 0 || device >= n) {
    fail(nullptr, IRP_ERR_BAD_ARG, "device %d out of range (%d devices)", device, n);
    return nullptr;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
    fail(nullptr, IRP_ERR_NO_DEVICE, "device %d is sm_%d%d; this build targets sm_100a (B200)", device, prop.major, prop.minor);
    return nullptr;
  }
  ctx = new irp_ctx();
  ctx->icc_profiles.push_back(std::make_shared<const std::vector<uint8_t>>());                        // 0: default, none
  ctx->icc_profiles.push_back(std::make_shared<const std::vector<uint8_t>>(make_srgb_profile()));     // IRP_ICC_SRGB
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (opts) memcpy(&ctx->opts, opts, std::min<size_t>(opts->struct_size, sizeof(irp_opts)));
  auto bail = [&](const char* what, cudaError_t err) {
    fail(nullptr, IRP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
    irp_destroy(ctx);
    return (irp_ctx*)nullptr;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  ctx->stream = ctx->own_stream;
  if ((e = cudaStreamCreateWithFlags(&ctx->copy_in_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->copy_out_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if (ctx->opts.staging_bytes) ctx->chunk_bytes = ctx->opts.staging_bytes;
  for (auto& ev : ctx->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  ClassifyTables* ht = new ClassifyTables();
  const bool cie = ctx->opts.luma_mode == IRP_LUMA_CIE;
  memcpy(ht->lut[0], cie ? kGreyLut1_r : kGreyLut0_r, sizeof ht->lut[0]);
  memcpy(ht->lut[1], cie ? kGreyLut1_g : kGreyLut0_g, sizeof ht->lut[1]);
  memcpy(ht->lut[2], cie ? kGreyLut1_b : kGreyLut0_b, sizeof ht->lut[2]);
  memcpy(ht->inv, cie ? kGreyInv1 : kGreyInv0, sizeof ht->inv);
  if ((e = cudaMalloc(&ctx->d_tables, sizeof(ClassifyTables))) != cudaSuccess) { delete ht; return bail("cudaMalloc(tables)", e); }
  e = cudaMemcpy(ctx->d_tables, ht, sizeof(ClassifyTables), cudaMemcpyHostToDevice);
  delete ht;
  if (e != cudaSuccess) return bail("cudaMemcpy(tables)", e);
  if ((e = cudaMalloc(&ctx->d_error_flag, 2 * sizeof(int))) != cudaSuccess) return bail("cudaMalloc(flag)", e);
  if ((e = cudaMemset(ctx->d_error_flag, 0, 2 * sizeof(int))) != cudaSuccess) return bail("cudaMemset(flag)", e);
  if ((e = cudaMallocHost(&ctx->h_error_flag, 2 * sizeof(int))) != cudaSuccess) return bail("cudaMallocHost(flag)", e);
  ctx->h_error_flag[0] = ctx->h_error_flag[1] = 0;
  // where does dynamic shared memory start in the .shared window? (the classify kernel aligns its
  // lookup tables to that address space and has no static shared memory of its own)
  probe_smem_base_kernel<<<1, 32, 1024, ctx->own_stream>>>(reinterpret_cast<uint32_t*>(ctx->d_error_flag + 1));
  if ((e = cudaMemcpyAsync(ctx->h_error_flag + 1, ctx->d_error_flag + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->own_stream)) != cudaSuccess)
    return bail("probe copy", e);
  if ((e = cudaStreamSynchronize(ctx->own_stream)) != cudaSuccess) return bail("probe_smem_base_kernel", e);
  ctx->smem_base = (uint32_t)ctx->h_error_flag[1];
  // The dynamic shared-memory attribute is per function and process-wide: raise it once to the
  // device's opt-in limit instead of tracking per-context maxima.
  ctx->smem_optin = prop.sharedMemPerBlockOptin - 4096;  // leave room for the kernels' small static arrays
  const void* kernels[6] = {(const void*)classify_kernel<1>, (const void*)classify_kernel<3>, (const void*)classify_kernel<4>,
                            (const void*)resize_kernel<1>,   (const void*)resize_kernel<3>,   (const void*)resize_kernel<4>};
  for (const void* k : kernels)
    if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin)) != cudaSuccess)
      return bail("cudaFuncSetAttribute(max dynamic shared memory)", e);
  {  // the streaming kernel has no static shared memory and takes the whole opt-in limit
    const BulkMap bm = make_bulk_map(ctx->smem_base);
    const char* off = getenv("IRP_NO_BULK");
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ctx->encode_tiled, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      ctx->encode_tiled = nullptr;
    ctx->bulk_ok = ctx->encode_tiled && bm.end - ctx->smem_base <= prop.sharedMemPerBlockOptin && !(off && off[0] == '1');
    ctx->smem_optin_full = prop.sharedMemPerBlockOptin;
    const char* off2 = getenv("IRP_NO_RTMA");
    ctx->rtma_ok = !(off2 && off2[0] == '1');
    if (ctx->bulk_ok && (e = cudaFuncSetAttribute((const void*)resize_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)prop.sharedMemPerBlockOptin - 2048)) != cudaSuccess)
      return bail("cudaFuncSetAttribute(resize_tma_kernel)", e);
    const char* off3 = getenv("IRP_NO_RMMA");
    ctx->rmma_ok = ctx->bulk_ok && !(off3 && off3[0] == '1');
    if (ctx->rmma_ok && (e = cudaFuncSetAttribute((const void*)resize_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)prop.sharedMemPerBlockOptin - 1024)) != cudaSuccess)
      return bail("cudaFuncSetAttribute(resize_mma_kernel)", e);
    if (ctx->bulk_ok &&
        (e = cudaFuncSetAttribute((const void*)classify_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)prop.sharedMemPerBlockOptin)) != cudaSuccess)
      return bail("cudaFuncSetAttribute(classify_bulk_kernel)", e);
  }
  return ctx;
}

void irp_destroy(irp_ctx* ctx) {
  if (!ctx) return;
  // the dispatcher drains its queue on stop and may still be inside a lane-split batch: stop and join it
  // first, then take the lanes away (under their lock), then the context itself (under its lock)
  {
    std::lock_guard<std::mutex> lk(ctx->qmu);
    ctx->stop = true;
  }
  ctx->qcv.notify_all();
  if (ctx->dispatcher_started && ctx->dispatcher.joinable()) ctx->dispatcher.join();
  {
    std::vector<irp_ctx*> lanes;
    {
      std::lock_guard<std::mutex> lk(ctx->lanes_mu);
      lanes.swap(ctx->lanes);
    }
    for (irp_ctx* l : lanes) irp_destroy(l);
  }
  { std::lock_guard<std::mutex> lk(ctx->mu); }   // a synchronous call still running on another thread finishes first
  cudaSetDevice(ctx->device);
  if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
  for (auto& ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
  for (auto& ev : ctx->sync_events) cudaEventDestroy(ev);
  for (auto& ev : ctx->timing_events) cudaEventDestroy(ev);
  if (ctx->copy_in_stream) cudaStreamDestroy(ctx->copy_in_stream);
  if (ctx->copy_out_stream) cudaStreamDestroy(ctx->copy_out_stream);
  for (DevBuf* b : {&ctx->d_desc, &ctx->d_acc, &ctx->d_stage_in, &ctx->d_stage_out, &ctx->d_orient, &ctx->d_jobs, &ctx->d_tmaps, &ctx->d_rtjobs, &ctx->d_rtmaps,
                    &ctx->d_jdata, &ctx->d_jmeta, &ctx->d_jstate, &ctx->d_jcoef, &ctx->d_jplane, &ctx->d_jpix, &ctx->d_emeta,
                    &ctx->d_eblk, &ctx->d_ebits, &ctx->d_eout, &ctx->d_epix, &ctx->d_ehuff, &ctx->d_mmjobs, &ctx->d_mmmaps, &ctx->d_mmctr})
    b->release();
  for (PinBuf* b : {&ctx->h_desc, &ctx->h_acc, &ctx->h_jobs, &ctx->h_tmaps, &ctx->h_rtjobs, &ctx->h_rtmaps, &ctx->h_jdata, &ctx->h_jmeta, &ctx->h_emeta, &ctx->h_ehuff, &ctx->h_mmjobs, &ctx->h_mmmaps}) b->release();
  for (void* p : ctx->plan_chunks) cudaFree(p);
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  if (ctx->d_error_flag) cudaFree(ctx->d_error_flag);
  if (ctx->h_error_flag) cudaFreeHost(ctx->h_error_flag);
  for (void* p : ctx->pinned) cudaFreeHost(p);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

const char* irp_last_error(const irp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int irp_set_stream(irp_ctx* ctx, void* cuda_stream) {
  if (!ctx) return IRP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return IRP_OK;
}

int irp_get_timing(const irp_ctx* ctx, irp_timing* out) {
  if (!ctx || !out) return IRP_ERR_BAD_ARG;
  *out = ctx->timing;
  return IRP_OK;
}

int irp_preprocess_dims(int width, int height, int exif_orientation, int* out_w, int* out_h) {
  if (width <= 0 || height <= 0 || !out_w || !out_h) return IRP_ERR_BAD_ARG;
  int o = (exif_orientation >= 1 && exif_orientation <= 8) ? exif_orientation : 1;
  double f;
  preprocess_dims(width, height, o, out_w, out_h, &f);
  return IRP_OK;
}

int irp_fusion_dims(int width, int height, int exif_orientation, int* out_w, int* out_h, int* off_x, int* off_y) {
  if (width <= 0 || height <= 0 || !out_w || !out_h || !off_x || !off_y) return IRP_ERR_BAD_ARG;
  int o = (exif_orientation >= 1 && exif_orientation <= 8) ? exif_orientation : 1;
  double f;
  fusion_dims(width, height, o, out_w, out_h, off_x, off_y, &f);
  return IRP_OK;
}

int irp_scores_from_moments(irp_result* r, int width, int height, int channels, int is_jpeg) {
  if (!r || width <= 0 || height <= 0 || channels < 1 || channels > 4) return IRP_ERR_BAD_ARG;
  const uint64_t N = (uint64_t)width * height;
  const int C = channels;
  double mean[4], stdev[4];
  uint64_t o_sum = 0, o_sumsq = 0;
  for (int ch = 0; ch < C; ch++) {
    double s = (double)r->sum[ch], s2 = (double)r->sumsq[ch], vals = (double)N;
    mean[ch] = s / vals;                                        // vips_stats mean
    stdev[ch] = std::sqrt(std::fabs(s2 - (s * s / vals)) / (vals - 1));  // vips_stats deviation
    o_sum += r->sum[ch];
    o_sumsq += r->sumsq[ch];
  }
  // classifier.js:118-121
  r->score[IRP_SCORE_BLUR] = jsmax(0, 1.0 - jsmin(pop_variance(r->e_sum[0], r->e_sumsq[0], N) / IRP_BLUR_VAR_DIVISOR, 1.0));
  // classifier.js:145-146
  r->score[IRP_SCORE_NOISE] = jsmin(std::sqrt(pop_variance(r->e_sum[1], r->e_sumsq[1], N)) / IRP_NOISE_STD_DIVISOR, 1.0);
  // classifier.js:159-167
  {
    double sum = 0;
    for (int ch = 0; ch < C; ch++) sum = sum + mean[ch];
    double nb = (sum / C) / 255;
    r->score[IRP_SCORE_LOWLIGHT] = nb < IRP_LOWLIGHT_KNEE ? jsmin((IRP_LOWLIGHT_KNEE - nb) * 2, 1.0) : 0.0;
  }
  // classifier.js:180-186,296-303
  if (is_jpeg) {
    double delta = jsmax(0, pop_variance(o_sum, o_sumsq, N * C) - pop_variance(r->b_sum, r->b_sumsq, N * C));
    r->score[IRP_SCORE_COMPRESSION] = jsmin(jsmin(delta / IRP_COMPRESSION_DIVISOR, 1.0), 1.0);
  } else {
    r->score[IRP_SCORE_COMPRESSION] = 0.0;
  }
  // classifier.js:335-336
  r->score[IRP_SCORE_SCRATCH] = jsmin(jsmin(((double)r->scratch_v + (double)r->scratch_h) / IRP_SCRATCH_DIVISOR, 1.0), 1.0);
  // classifier.js:223-228,272-286
  {
    double colorfulness = C < 3 ? 0.5
                                : jsmin(std::sqrt(std::pow(stdev[0], 2) + std::pow(stdev[1], 2) + std::pow(stdev[2], 2)) / 255, 1.0);
    double sum = 0;
    for (int ch = 0; ch < C; ch++) sum = sum + stdev[ch];
    double contrast = jsmin((sum / C) / IRP_CONTRAST_DIVISOR, 1.0);
    r->score[IRP_SCORE_FADE] = jsmin((1.0 - colorfulness) * 0.6 + (1.0 - contrast) * 0.4, 1.0);
  }
  // classifier.js:240-253
  if (C < 3) {
    r->score[IRP_SCORE_COLORSHIFT] = 0.0;
  } else {
    double avg = (mean[0] + mean[1] + mean[2]) / 3;
    double rd = avg > 0 ? std::fabs(mean[0] - avg) / avg : 0, gd = avg > 0 ? std::fabs(mean[1] - avg) / avg : 0,
           bd = avg > 0 ? std::fabs(mean[2] - avg) / avg : 0;
    r->score[IRP_SCORE_COLORSHIFT] = jsmin(jsmax(jsmax(rd, gd), bd) * 2, 1.0);
  }
  return irp_top_issues(r);
}

// PromptEnhancerService._identifyTopIssues / _determineSeverity (promptEnhancer.js:121-145)
int irp_top_issues(irp_result* r) {
  if (!r) return IRP_ERR_BAD_ARG;
  int idx[IRP_NUM_SCORES], n = 0;
  for (int k = 0; k < IRP_NUM_SCORES; k++)
    if (r->score[k] > IRP_ISSUE_THRESHOLD) idx[n++] = k;
  std::stable_sort(idx, idx + n, [&](int a, int b) { return r->score[a] > r->score[b]; });
  n = std::min(n, 3);
  for (int k = 0; k < 3; k++) {
    if (k >= n) {
      r->issues[k] = IRP_NO_ISSUE;
      continue;
    }
    const double c = r->score[idx[k]];
    r->issues[k] = IRP_ISSUE(c >= 0.7 ? IRP_SEVERITY_HIGH : (c >= 0.5 ? IRP_SEVERITY_MEDIUM : IRP_SEVERITY_LOW), idx[k]);
  }
  r->issues[3] = (uint8_t)n;
  return IRP_OK;
}

int irp_grey_tables(int luma_mode, uint32_t lut_r[256], uint32_t lut_g[256], uint32_t lut_b[256], uint32_t inv[4096]) {
  if (!lut_r || !lut_g || !lut_b || !inv) return IRP_ERR_BAD_ARG;
  const bool cie = luma_mode == IRP_LUMA_CIE;
  memcpy(lut_r, cie ? kGreyLut1_r : kGreyLut0_r, 1024);
  memcpy(lut_g, cie ? kGreyLut1_g : kGreyLut0_g, 1024);
  memcpy(lut_b, cie ? kGreyLut1_b : kGreyLut0_b, 1024);
  memcpy(inv, cie ? kGreyInv1 : kGreyInv0, 4096 * 4);
  return IRP_OK;
}

int irp_classify_batch(irp_ctx* ctx, const irp_image_desc* imgs, int n, irp_result* results) {
  if (!results) return ctx ? fail(ctx, IRP_ERR_BAD_ARG, "null results") : IRP_ERR_BAD_ARG;
  return run_batch(ctx, imgs, n, results, nullptr, 0);
}
int irp_preprocess_batch(irp_ctx* ctx, const irp_image_desc* imgs, int n, irp_out_desc* outs) {
  if (!outs) return ctx ? fail(ctx, IRP_ERR_BAD_ARG, "null outputs") : IRP_ERR_BAD_ARG;
  return run_batch(ctx, imgs, n, nullptr, outs, 0);
}
int irp_analyze_batch(irp_ctx* ctx, const irp_image_desc* imgs, int n, irp_result* results, irp_out_desc* outs) {
  if (!results || !outs) return ctx ? fail(ctx, IRP_ERR_BAD_ARG, "null results/outputs") : IRP_ERR_BAD_ARG;
  return run_batch(ctx, imgs, n, results, outs, 0);
}
int irp_fusion_prepare_batch(irp_ctx* ctx, const irp_image_desc* imgs, int n_groups, irp_out_desc* canvases) {
  if (!canvases) return ctx ? fail(ctx, IRP_ERR_BAD_ARG, "null canvases") : IRP_ERR_BAD_ARG;
  return run_batch(ctx, imgs, n_groups * IRP_FUSION_MAX_IMAGES, nullptr, canvases, 1);
}

// ---- compressed input: baseline JPEG decoded on the device ----
int irp_jpeg_info(const uint8_t* data, size_t size, int* width, int* height, int* channels) {
  if (!data || !width || !height || !channels) return IRP_ERR_BAD_ARG;
  HostJpeg J;
  const char* why = "";
  const int rc = jpeg_parse(data, size, &J, &why);
  if (rc) return rc;
  *width = J.w;
  *height = J.h;
  *channels = J.ncomp == 1 ? 1 : 3;
  return IRP_OK;
}

int irp_decode_jpeg_batch(irp_ctx* ctx, const irp_jpeg_desc* jpegs, int n, irp_out_desc* outs) {
  if (!ctx || n < 0 || (n && (!jpegs || !outs))) return IRP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->err.clear();
  if (!n) return IRP_OK;
  ctx->timing = irp_timing{};
  std::vector<JpegPlaced> pl;
  ctx->jpeg_allow_scale = true;
  int rc = decode_jpegs_locked(ctx, jpegs, n, &pl);
  ctx->jpeg_allow_scale = false;
  if (rc) return rc;
  for (int i = 0; i < n; i++) {   // every output is checked before the first copy is queued
    const irp_out_desc& od = outs[i];
    const size_t tight = (size_t)pl[i].w * pl[i].c, pitch = od.pitch ? od.pitch : tight;
    if (!od.pixels || pitch < tight || od.capacity < pitch * (size_t)(pl[i].h - 1) + tight) {
      drain_streams(ctx);
      return fail(ctx, IRP_ERR_CAPACITY, "output %d: capacity %zu too small for %dx%dx%d", i, od.capacity, pl[i].w, pl[i].h, pl[i].c);
    }
  }
  for (int i = 0; i < n; i++) {
    irp_out_desc& od = outs[i];
    const size_t tight = (size_t)pl[i].w * pl[i].c, pitch = od.pitch ? od.pitch : tight;
    od.width = pl[i].w;
    od.height = pl[i].h;
    od.channels = pl[i].c;
    CK(cudaMemcpy2DAsync(od.pixels, pitch, pl[i].px, pl[i].pitch, tight, pl[i].h, od.on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                         ctx->stream));
  }
  CK(wait_stream(ctx, ctx->stream));
  return IRP_OK;
}

// Cut [0, n) into the context's lanes and run f(lane context, begin, count) on each from its own host thread
// (lane 0 is the context itself, on the calling thread).  At least 8 files per lane; IRP_LANES overrides the
// default of 8 (1 = off).  The first failing lane's status and message are returned.
static int run_in_lanes(irp_ctx* ctx, int n, const std::function<int(irp_ctx*, int, int)>& f) {
  // as many lanes as this process has cores to drive them with (8 on a box of its own, 4 when eight ranks share 32 cores)
  static const int env_lanes = getenv("IRP_LANES") ? atoi(getenv("IRP_LANES")) : std::max(2, std::min(8, host_core_share()));
  static const int env_min = getenv("IRP_LANE_MIN") ? std::max(1, atoi(getenv("IRP_LANE_MIN"))) : 8;
  const int want = ctx->is_lane ? 1 : std::max(1, std::min(env_lanes, n / env_min));
  if (want == 1) return f(ctx, 0, n);
  {
    std::lock_guard<std::mutex> lk(ctx->lanes_mu);
    while ((int)ctx->lanes.size() < want - 1) {
      irp_ctx* l = irp_create(ctx->device, &ctx->opts);
      if (!l) return f(ctx, 0, n);      // no memory for another lane: run unsplit
      l->is_lane = true;
      l->parent = ctx;      // profiles (IRP_JPEG_ICC) resolve through the parent
      ctx->lanes.push_back(l);
    }
  }
  std::vector<irp_ctx*> lanes;
  {
    std::lock_guard<std::mutex> lk(ctx->lanes_mu);
    lanes = ctx->lanes;
  }
  std::vector<int> rc(want, IRP_OK);
  std::vector<std::thread> th;
  auto part = [&](int k) { return (int)((long long)n * k / want); };
  for (int k = 1; k < want; k++) th.emplace_back([&, k] { rc[k] = f(lanes[k - 1], part(k), part(k + 1) - part(k)); });
  rc[0] = f(ctx, 0, part(1));
  for (auto& t : th) t.join();
  for (int k = 1; k < want; k++) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->timing.kernel_launches += lanes[k - 1]->timing.kernel_launches;
    ctx->timing.classify_ms += lanes[k - 1]->timing.classify_ms;       // sums over lanes (they overlap in time)
    ctx->timing.preprocess_ms += lanes[k - 1]->timing.preprocess_ms;
    if (rc[k] != IRP_OK && rc[0] == IRP_OK) {
      rc[0] = rc[k];
      ctx->err = lanes[k - 1]->err;
    }
  }
  return rc[0];
}

static int analyze_jpeg_single(irp_ctx* ctx, const irp_jpeg_desc* jpegs, int n, irp_result* results, irp_out_desc* outs);
int irp_analyze_jpeg_batch(irp_ctx* ctx, const irp_jpeg_desc* jpegs, int n, irp_result* results, irp_out_desc* outs) {
  if (!ctx || n < 0 || (n && (!jpegs || (!results && !outs)))) return IRP_ERR_BAD_ARG;
  return run_in_lanes(ctx, n, [&](irp_ctx* lane, int b, int cnt) {
    return analyze_jpeg_single(lane, jpegs + b, cnt, results ? results + b : nullptr, outs ? outs + b : nullptr);
  });
}
static int analyze_jpeg_single(irp_ctx* ctx, const irp_jpeg_desc* jpegs, int n, irp_result* results, irp_out_desc* outs) {
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->err.clear();
  if (!n) return IRP_OK;
  ctx->timing = irp_timing{};
  cudaEventRecord(ctx->ev[1], ctx->stream);
  std::vector<JpegPlaced> pl;
  int rc = decode_jpegs_locked(ctx, jpegs, n, &pl);
  if (rc) return rc;
  cudaEventRecord(ctx->ev[2], ctx->stream);
  const uint32_t decode_launches = ctx->timing.kernel_launches;
  std::vector<irp_image_desc> descs(n);
  for (int i = 0; i < n; i++)
    descs[i] = irp_image_desc{pl[i].px, pl[i].pitch, pl[i].w, pl[i].h, pl[i].c, 1, jpegs[i].exif_orientation, 1};
  rc = run_batch_locked(ctx, descs.data(), n, results, outs, 0);
  ctx->timing.kernel_launches += decode_launches;
  float ms = 0;
  if (rc == IRP_OK && cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]) == cudaSuccess) ctx->timing.h2d_ms = ms;   // upload of the compressed bytes + device decode
  return rc;
}

// ---- compressed output: baseline JPEG encoded on the device ----
int irp_set_output_icc(irp_ctx* ctx, const uint8_t* profile, size_t size) {
  if (!ctx || (size && !profile) || size > 255u * 65519u) return IRP_ERR_BAD_ARG;
  auto prof = std::make_shared<const std::vector<uint8_t>>(profile, profile + size);
  std::lock_guard<std::mutex> lock(ctx->icc_mu);
  ctx->icc_profiles[0] = prof;   // calls already running keep the profile they resolved
  return IRP_OK;
}

int irp_register_icc(irp_ctx* ctx, const uint8_t* profile, size_t size) {
  if (!ctx || !profile || !size || size > 255u * 65519u) return IRP_ERR_BAD_ARG;
  auto prof = std::make_shared<const std::vector<uint8_t>>(profile, profile + size);
  std::lock_guard<std::mutex> lock(ctx->icc_mu);
  for (size_t k = 1; k < ctx->icc_profiles.size(); k++)
    if (*ctx->icc_profiles[k] == *prof) return (int)k;   // registering the same bytes again returns the same id
  if (ctx->icc_profiles.size() >= 256) return IRP_ERR_CAPACITY;
  ctx->icc_profiles.push_back(prof);
  return (int)ctx->icc_profiles.size() - 1;
}

int irp_get_icc(irp_ctx* ctx, int id, uint8_t* out, size_t capacity, size_t* size) {
  if (!size || (!ctx && id != IRP_ICC_SRGB)) return IRP_ERR_BAD_ARG;
  // the generated sRGB profile needs no context (and no device): a host-only helper like irp_preprocess_dims
  auto prof = ctx ? resolve_icc(ctx, id) : std::make_shared<const std::vector<uint8_t>>(make_srgb_profile());
  if (!prof) return IRP_ERR_BAD_ARG;
  *size = prof->size();
  if (out && capacity >= prof->size() && !prof->empty()) memcpy(out, prof->data(), prof->size());
  return (out && capacity < prof->size()) ? IRP_ERR_CAPACITY : IRP_OK;
}

int irp_encode_jpeg_batch(irp_ctx* ctx, const irp_image_desc* imgs, int n, int quality, irp_jpeg_out* outs) {
  if (!ctx || n < 0 || (n && (!imgs || !outs))) return IRP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->err.clear();
  if (!n) return IRP_OK;
  ctx->timing = irp_timing{};
  CK(cudaSetDevice(ctx->device));
  // host-resident sources are staged once (an encode-only call is a test / tooling entry; the serving paths below
  // encode what the resize kernels left in HBM)
  std::vector<EncSrc> src(n);
  size_t stage = 0;
  for (int i = 0; i < n; i++) {
    if (!imgs[i].pixels || imgs[i].width < 1 || imgs[i].height < 1 || (imgs[i].channels != 1 && imgs[i].channels != 3) ||
        imgs[i].pitch < (size_t)imgs[i].width * imgs[i].channels)
      return fail(ctx, IRP_ERR_BAD_ARG, "encode %d: bad image (1 or 3 channels of u8)", i);
    if (!imgs[i].on_device) stage += round_up(round_up((size_t)imgs[i].width * imgs[i].channels, 16) * imgs[i].height, 256);
  }
  CK(ctx->d_epix.reserve(stage + 256));
  size_t off = 0;
  for (int i = 0; i < n; i++) {
    const irp_image_desc& d = imgs[i];
    if (d.on_device) {
      src[i] = EncSrc{d.pixels, d.pitch, d.width, d.height, d.channels};
      continue;
    }
    const size_t tight = (size_t)d.width * d.channels, pitch = round_up(tight, 16);
    uint8_t* dst = (uint8_t*)ctx->d_epix.p + off;
    CK(cudaMemcpy2DAsync(dst, pitch, d.pixels, d.pitch, tight, d.height, cudaMemcpyHostToDevice, ctx->stream));
    src[i] = EncSrc{dst, pitch, d.width, d.height, d.channels};
    off += round_up(pitch * d.height, 256);
  }
  return encode_images_locked(ctx, src.data(), n, quality, outs);
}

// preprocess into context-owned HBM, then encode from there
static int preprocess_encode_locked(irp_ctx* ctx, const irp_image_desc* descs, int n, irp_result* results, int quality, irp_jpeg_out* outs) {
  std::vector<irp_out_desc> od(n);
  std::vector<EncSrc> src(n);
  size_t total = 0;
  for (int i = 0; i < n; i++) {
    int ow = 0, oh = 0;
    if (irp_preprocess_dims(descs[i].width, descs[i].height, descs[i].exif_orientation, &ow, &oh) != IRP_OK)
      return fail(ctx, IRP_ERR_BAD_ARG, "image %d: bad dimensions %dx%d", i, descs[i].width, descs[i].height);
    const int oc = descs[i].channels == 1 ? 1 : 3;
    const size_t pitch = round_up((size_t)ow * oc, 16);
    od[i] = irp_out_desc{nullptr, pitch, pitch * oh, 0, 0, 0, 1};
    src[i] = EncSrc{nullptr, pitch, ow, oh, oc};
    total += round_up(pitch * oh, 256);
  }
  CK(ctx->d_epix.reserve(total + 256));
  size_t off = 0;
  for (int i = 0; i < n; i++) {
    od[i].pixels = (uint8_t*)ctx->d_epix.p + off;
    src[i].px = od[i].pixels;
    off += round_up(od[i].capacity, 256);
  }
  const auto t0 = std::chrono::steady_clock::now();
  int rc = run_batch_locked(ctx, descs, n, results, od.data(), 0);
  if (rc) return rc;
  if (getenv("IRP_TRACE"))
    fprintf(stderr, "transcode: classify + preprocess returned after %.2f ms (includes waiting for the decode kernels)\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  for (int i = 0; i < n; i++)
    if (od[i].width != src[i].w || od[i].height != src[i].h || od[i].channels != src[i].c)
      return fail(ctx, IRP_ERR_UNSUPPORTED, "image %d: preprocess wrote %dx%dx%d, expected %dx%dx%d", i, od[i].width, od[i].height, od[i].channels, src[i].w,
                  src[i].h, src[i].c);
  const irp_timing t = ctx->timing;
  rc = encode_images_locked(ctx, src.data(), n, quality, outs);
  const uint32_t enc_launches = ctx->timing.kernel_launches - t.kernel_launches;
  ctx->timing = t;
  ctx->timing.kernel_launches += enc_launches;
  return rc;
}

int irp_analyze_encode_batch(irp_ctx* ctx, const irp_image_desc* imgs, int n, irp_result* results, int quality, irp_jpeg_out* outs) {
  if (!ctx || n < 0 || (n && (!imgs || !outs))) return IRP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->err.clear();
  if (!n) return IRP_OK;
  return preprocess_encode_locked(ctx, imgs, n, results, quality, outs);
}

static int transcode_single(irp_ctx* ctx, const irp_jpeg_desc* jpegs, int n, irp_result* results, int quality, irp_jpeg_out* outs);
int irp_transcode_jpeg_batch(irp_ctx* ctx, const irp_jpeg_desc* jpegs, int n, irp_result* results, int quality, irp_jpeg_out* outs) {
  if (!ctx || n < 0 || (n && (!jpegs || !outs))) return IRP_ERR_BAD_ARG;
  return run_in_lanes(ctx, n, [&](irp_ctx* lane, int b, int cnt) {
    return transcode_single(lane, jpegs + b, cnt, results ? results + b : nullptr, quality, outs + b);
  });
}
static int transcode_single(irp_ctx* ctx, const irp_jpeg_desc* jpegs, int n, irp_result* results, int quality, irp_jpeg_out* outs) {
  std::lock_guard<std::mutex> lock(ctx->mu);
  ctx->err.clear();
  if (!n) return IRP_OK;
  ctx->timing = irp_timing{};
  const bool trace = getenv("IRP_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto since = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  std::vector<JpegPlaced> pl;
  int rc = decode_jpegs_locked(ctx, jpegs, n, &pl);
  if (rc) return rc;
  const double t_dec = since();
  const uint32_t decode_launches = ctx->timing.kernel_launches;
  std::vector<irp_image_desc> descs(n);
  for (int i = 0; i < n; i++)
    descs[i] = irp_image_desc{pl[i].px, pl[i].pitch, pl[i].w, pl[i].h, pl[i].c, 1, jpegs[i].exif_orientation, 1};
  rc = preprocess_encode_locked(ctx, descs.data(), n, results, quality, outs);
  ctx->timing.kernel_launches += decode_launches;
  if (trace) fprintf(stderr, "transcode: %d files, host wall: decode enqueued at %.2f ms, call done at %.2f ms\n", n, t_dec, since());
  return rc;
}

// ---- concurrent single-image requests ----
static void run_transcode_requests(irp_ctx* ctx, std::vector<irp_request*>& sel) {
  // one batched submission per quality present; a failing batch is retried request by request
  while (!sel.empty()) {
    const int quality = sel[0]->quality;
    std::vector<irp_request*> grp, rest;
    for (irp_request* r : sel) (r->quality == quality ? grp : rest).push_back(r);
    sel.swap(rest);
    const int n = (int)grp.size();
    std::vector<irp_jpeg_desc> jd(n);
    std::vector<irp_result> results(n);
    std::vector<irp_jpeg_out> outs(n);
    for (int i = 0; i < n; i++) {
      jd[i] = grp[i]->jpeg;
      outs[i] = *grp[i]->enc;
    }
    int rc = irp_transcode_jpeg_batch(ctx, jd.data(), n, results.data(), quality, outs.data());
    for (int i = 0; i < n; i++) {
      if (rc != IRP_OK && n > 1) {
        outs[i] = *grp[i]->enc;
        grp[i]->status = irp_transcode_jpeg_batch(ctx, &jd[i], 1, &results[i], quality, &outs[i]);
      } else {
        grp[i]->status = rc;
      }
      if (grp[i]->status != IRP_OK) grp[i]->err = ctx->err;
      if (grp[i]->status == IRP_OK || grp[i]->status == IRP_ERR_CAPACITY) *grp[i]->enc = outs[i];   // size needed on a capacity error
      if (grp[i]->status == IRP_OK && grp[i]->result) *grp[i]->result = results[i];
    }
  }
}

static void run_requests(irp_ctx* ctx, std::vector<irp_request*>& all) {
  std::vector<irp_request*> reqs, trans;
  for (irp_request* r : all) (r->enc ? trans : reqs).push_back(r);
  run_transcode_requests(ctx, trans);
  // raw-pixel and JPEG-file requests, each as classify+preprocess / classify / preprocess: every kind present
  // in the queue is one batched submission
  for (int kind = 0; kind < 6; kind++) {
    const bool jpeg_kind = kind >= 3;
    std::vector<irp_request*> sel;
    for (irp_request* r : reqs) {
      const int k = ((r->result && r->out) ? 0 : (r->result ? 1 : 2)) + (r->is_jpeg_file ? 3 : 0);
      if (k == kind) sel.push_back(r);
    }
    if (sel.empty()) continue;
    const int n = (int)sel.size();
    std::vector<irp_image_desc> descs(n);
    std::vector<irp_jpeg_desc> jdescs(n);
    std::vector<irp_result> results(n);
    std::vector<irp_out_desc> outs(n);
    for (int i = 0; i < n; i++) {
      descs[i] = sel[i]->img;
      jdescs[i] = sel[i]->jpeg;
      if (sel[i]->out) outs[i] = *sel[i]->out;
    }
    const int k3 = kind % 3;
    auto call = [&](int b, int cnt) {
      irp_result* rp = k3 == 2 ? nullptr : results.data() + b;
      irp_out_desc* op = k3 == 1 ? nullptr : outs.data() + b;
      return jpeg_kind ? irp_analyze_jpeg_batch(ctx, jdescs.data() + b, cnt, rp, op) : run_batch(ctx, descs.data() + b, cnt, rp, op, 0);
    };
    int rc = call(0, n);
    if (rc == IRP_OK) {
      for (int i = 0; i < n; i++) sel[i]->status = IRP_OK;
    } else if (n == 1) {
      sel[0]->status = rc;
      sel[0]->err = ctx->err;
    } else {  // isolate the request(s) that fail: one bad upload must not fail its batch-mates
      for (int i = 0; i < n; i++) {
        sel[i]->status = call(i, 1);
        if (sel[i]->status != IRP_OK) sel[i]->err = ctx->err;
      }
    }
    for (int i = 0; i < n; i++) {
      if (sel[i]->status != IRP_OK) continue;
      if (sel[i]->result) *sel[i]->result = results[i];
      if (sel[i]->out) *sel[i]->out = outs[i];
    }
  }
}

static void dispatcher_main(irp_ctx* ctx) {
  const size_t max_batch = 64;
  for (;;) {
    std::vector<irp_request*> reqs;
    {
      std::unique_lock<std::mutex> lk(ctx->qmu);
      ctx->qcv.wait(lk, [&] { return ctx->stop || !ctx->queue.empty(); });
      if (ctx->queue.empty()) return;  // stop requested and nothing left
      if (ctx->queue.size() < max_batch && !ctx->stop)  // a short window for concurrent callers to join the batch
        ctx->qcv.wait_for(lk, std::chrono::microseconds(100), [&] { return ctx->stop || ctx->queue.size() >= max_batch; });
      while (!ctx->queue.empty() && reqs.size() < max_batch) {
        reqs.push_back(ctx->queue.front());
        ctx->queue.pop_front();
      }
    }
    run_requests(ctx, reqs);
    {
      std::lock_guard<std::mutex> lk(ctx->qmu);
      for (irp_request* r : reqs) r->done = true;
    }
    ctx->done_cv.notify_all();
  }
}

int irp_submit(irp_ctx* ctx, const irp_image_desc* img, irp_result* result, irp_out_desc* out, irp_ticket* ticket) {
  if (!ctx || !img || !ticket || (!result && !out)) return IRP_ERR_BAD_ARG;
  irp_request* r = new irp_request();
  r->img = *img;
  r->result = result;
  r->out = out;
  {
    std::lock_guard<std::mutex> lk(ctx->qmu);
    if (ctx->stop) {
      delete r;
      return IRP_ERR_BAD_ARG;
    }
    if (!ctx->dispatcher_started) {
      ctx->dispatcher = std::thread(dispatcher_main, ctx);
      ctx->dispatcher_started = true;
    }
    ctx->queue.push_back(r);
  }
  ctx->qcv.notify_all();
  *ticket = r;
  return IRP_OK;
}

int irp_submit_jpeg(irp_ctx* ctx, const irp_jpeg_desc* jpeg, irp_result* result, irp_out_desc* out, irp_ticket* ticket) {
  if (!ctx || !jpeg || !jpeg->data || !ticket || (!result && !out)) return IRP_ERR_BAD_ARG;
  irp_request* r = new irp_request();
  r->jpeg = *jpeg;
  r->is_jpeg_file = true;
  r->result = result;
  r->out = out;
  {
    std::lock_guard<std::mutex> lk(ctx->qmu);
    if (ctx->stop) {
      delete r;
      return IRP_ERR_BAD_ARG;
    }
    if (!ctx->dispatcher_started) {
      ctx->dispatcher = std::thread(dispatcher_main, ctx);
      ctx->dispatcher_started = true;
    }
    ctx->queue.push_back(r);
  }
  ctx->qcv.notify_all();
  *ticket = r;
  return IRP_OK;
}

int irp_submit_transcode(irp_ctx* ctx, const irp_jpeg_desc* jpeg, irp_result* result, int quality, irp_jpeg_out* out, irp_ticket* ticket) {
  if (!ctx || !jpeg || !jpeg->data || !ticket || !out || !out->data) return IRP_ERR_BAD_ARG;
  irp_request* r = new irp_request();
  r->jpeg = *jpeg;
  r->is_jpeg_file = true;
  r->result = result;
  r->enc = out;
  r->quality = quality;
  {
    std::lock_guard<std::mutex> lk(ctx->qmu);
    if (ctx->stop) {
      delete r;
      return IRP_ERR_BAD_ARG;
    }
    if (!ctx->dispatcher_started) {
      ctx->dispatcher = std::thread(dispatcher_main, ctx);
      ctx->dispatcher_started = true;
    }
    ctx->queue.push_back(r);
  }
  ctx->qcv.notify_all();
  *ticket = r;
  return IRP_OK;
}

int irp_wait(irp_ctx* ctx, irp_ticket ticket, char* err, size_t err_capacity) {
  if (!ctx || !ticket) return IRP_ERR_BAD_ARG;
  irp_request* r = ticket;
  {
    std::unique_lock<std::mutex> lk(ctx->qmu);
    ctx->done_cv.wait(lk, [&] { return r->done; });
  }
  const int rc = r->status;
  if (err && err_capacity) {
    const size_t nb = std::min(err_capacity - 1, r->err.size());
    memcpy(err, r->err.data(), nb);
    err[nb] = 0;
  }
  delete r;
  return rc;
}

void* irp_dev_alloc(irp_ctx* ctx, size_t bytes) {
  if (!ctx) return nullptr;
  std::lock_guard<std::mutex> lock(ctx->mu);
  void* p = nullptr;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    fail(ctx, IRP_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}
int irp_dev_free(irp_ctx* ctx, void* p) {
  if (!ctx) return IRP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  CK(cudaFree(p));
  return IRP_OK;
}
void* irp_host_alloc_pinned(irp_ctx* ctx, size_t bytes) {
  if (!ctx) return nullptr;
  std::lock_guard<std::mutex> lock(ctx->mu);
  void* p = nullptr;
  cudaSetDevice(ctx->device);
  cudaError_t e = cudaMallocHost(&p, bytes);
  if (e != cudaSuccess) {
    fail(ctx, IRP_ERR_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  {
    std::lock_guard<std::mutex> lk(ctx->pin_mu);
    ctx->pinned.push_back(p);
  }
  return p;
}
int irp_host_free_pinned(irp_ctx* ctx, void* p) {
  if (!ctx) return IRP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  {
    std::lock_guard<std::mutex> lk(ctx->pin_mu);
    auto it = std::find(ctx->pinned.begin(), ctx->pinned.end(), p);
    if (it == ctx->pinned.end()) return fail(ctx, IRP_ERR_BAD_ARG, "irp_host_free_pinned: not an allocation of this context");
    ctx->pinned.erase(it);
  }
  CK(cudaFreeHost(p));
  return IRP_OK;
}
int irp_memcpy_h2d(irp_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx) return IRP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(wait_stream(ctx, ctx->stream));
  return IRP_OK;
}
int irp_memcpy_d2h(irp_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx) return IRP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(wait_stream(ctx, ctx->stream));
  return IRP_OK;
}
int irp_synchronize(irp_ctx* ctx) {
  if (!ctx) return IRP_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lock(ctx->mu);
  CK(cudaSetDevice(ctx->device));
  CK(wait_stream(ctx, ctx->stream));
  return IRP_OK;
}

}  // extern "C"
