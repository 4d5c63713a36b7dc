// umma_rate.cu — cycles per tcgen05.mma kind::i8 instruction for the shapes / operand layouts the resize kernel
// could use (M 128; N 64..256; A K-major or MN-major, with or without the 128-byte swizzle).  64 instructions are
// issued back to back on the same operands (contents irrelevant), then one commit; clock64 around issue + wait.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

struct Cfg { uint64_t a_desc, b_desc; uint32_t idesc, n_instr, kind_f16, issuers, lanes, same_d, per_commit; };

template <int PC>
__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem + (sbase - (uint32_t)__cvta_generic_to_shared(smem)))[i] = 0x01010101u;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_addr), "r"(c.issuers * (c.per_commit ? (c.n_instr / c.issuers) / c.per_commit : 1u)) : "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  long long t0 = 0, t1 = 0, t2 = 0;
  // issuers are lane 0 of the first `issuers` warps, or (lanes mode) the first `issuers` lanes of warp 0
  const bool is_issuer = c.lanes ? threadIdx.x < c.issuers : ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < c.issuers);
  const uint32_t my = c.lanes ? threadIdx.x : (threadIdx.x >> 5);
  if (is_issuer) {
    const uint64_t ad = c.a_desc + (uint64_t)((sbase >> 4) & 0x3FFF), bd = c.b_desc + (uint64_t)(((sbase + 96 * 1024) >> 4) & 0x3FFF);
    t0 = clock64();
    const uint32_t dcol = tm + (c.same_d ? 0u : my * 64u);
    if (PC == 0) {
      for (uint32_t k = 0; k < c.n_instr / c.issuers; k++) {
        if (c.kind_f16)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol), "l"(ad), "l"(bd), "r"(c.idesc) : "memory");
        else
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol), "l"(ad), "l"(bd), "r"(c.idesc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
    } else {
      for (uint32_t k = 0; k < c.n_instr / c.issuers / (PC ? PC : 1); k++) {
#pragma unroll
        for (int j = 0; j < PC; j++)
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(dcol), "l"(ad), "l"(bd), "r"(c.idesc) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
      }
    }
    t1 = clock64();
  }
  asm volatile("{\n.reg .pred q;\nW: mbarrier.try_wait.parity.shared::cta.b64 q, [%0], 0;\n@q bra D;\nbra W;\nD:\n}\n" ::"r"(bar_addr) : "memory");
  if (threadIdx.x == 0) {
    t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

static uint32_t idesc_i8(int M, int N, int a_mn, int b_mn) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
static uint32_t idesc_f16(int M, int N) {   // bf16 x bf16 -> f32, K-major
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
static uint64_t sdesc(uint32_t lbo, uint32_t sbo, int layout) {
  return ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

int main() {
  CK(cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(rate_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(rate_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long* d_out;
  CK(cudaMalloc(&d_out, 16));
  struct Case { const char* name; int M, N, a_kind, b_kind, f16; };   // kinds: 0 K-major none, 2 MN-major sw128, 1 MN-major none, 3 K-major sw128
  const Case cases[] = {
      {"i8 M128 N64  A MN sw128, B K none", 128, 64, 2, 0, 0},  {"i8 M128 N128 A MN sw128, B K none", 128, 128, 2, 0, 0},
      {"i8 M128 N256 A MN sw128, B K none", 128, 256, 2, 0, 0}, {"i8 M128 N64  A K none,  B K none", 128, 64, 0, 0, 0},
      {"i8 M128 N128 A K none,  B K none", 128, 128, 0, 0, 0},   {"i8 M128 N256 A K none,  B K none", 128, 256, 0, 0, 0},
      {"i8 M128 N256 A K sw128, B K sw128", 128, 256, 3, 3, 0},  {"i8 M128 N64  A MN none, B K none", 128, 64, 1, 0, 0},
      {"i8 M128 N256 A MN none, B K none", 128, 256, 1, 0, 0},   {"i8 M128 N64  A K sw128, B K sw128", 128, 64, 3, 3, 0},
      {"i8 M64  N64  A K none,  B K none", 64, 64, 0, 0, 0},     {"i8 M64  N256 A K none,  B K none", 64, 256, 0, 0, 0},
      {"bf16 M128 N256 K-major sw128", 128, 256, 3, 3, 1},        {"bf16 M128 N64 K-major sw128", 128, 64, 3, 3, 1},
  };
  for (const Case& c : cases) {
    Cfg g;
    auto mk = [&](int kind, int rows) -> uint64_t {
      switch (kind) {
        case 0: return sdesc(rows * 16, 128, 0);
        case 1: return sdesc(128, 64 * 16, 0);
        case 2: return sdesc(64 * 128, 1024, 2);
        default: return sdesc(16, 1024, 2);
      }
    };
    g.a_desc = mk(c.a_kind, c.M);
    g.b_desc = mk(c.b_kind, c.N);
    g.idesc = c.f16 ? idesc_f16(c.M, c.N) : idesc_i8(c.M, c.N, c.a_kind == 1 || c.a_kind == 2, c.b_kind == 1 || c.b_kind == 2);
    g.kind_f16 = c.f16;
    for (int mode = 0; mode < 3; mode++)
    for (int iss : {1, 2, 3, 4, 6, 8}) {
      const int n = 96;
      if (c.N * iss > 512 && mode != 2) continue;
      if (mode > 0 && iss == 1) continue;
      if (iss > 4 && mode == 0) continue;
      g.lanes = mode >= 1;
      g.same_d = mode == 2;
      g.issuers = iss;
      g.per_commit = 0;
      g.n_instr = n;
      rate_kernel<0><<<1, 128, 200 * 1024>>>(g, d_out);
      CK(cudaDeviceSynchronize());
      long long h[2];
      CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
      printf("%-36s %s issuers %d n=%2d: issue %6lld cycles, done %6lld cycles (%.1f per instruction)\n", c.name, mode == 0 ? "warps" : (mode == 1 ? "lanes" : "lanes, one accumulator"), iss, n, h[0], h[1], (double)h[1] / n);
    }
  }
  {  // what a commit costs: M128 N64, `lanes` lanes of one warp (or warps), a commit after every per_commit MMAs of a lane
    const Case c = {"i8 M128 N64 A MN sw128", 128, 64, 2, 0, 0};
    Cfg g;
    g.a_desc = sdesc(64 * 128, 1024, 2);
    g.b_desc = sdesc(64 * 16, 128, 0);
    g.idesc = idesc_i8(128, 64, 1, 0);
    g.kind_f16 = 0;
    g.same_d = 0;
    for (int lanes_mode = 0; lanes_mode < 2; lanes_mode++)
      for (int iss : {1, 2, 3, 4})
        for (int pc : {0, 1, 2, 3, 6}) {
          g.lanes = lanes_mode;
          g.issuers = iss;
          g.per_commit = pc;
          g.n_instr = 96 / iss / (pc ? pc : 1) * (pc ? pc : 1) * iss;
          switch (pc) {
            case 0: rate_kernel<0><<<1, 128, 200 * 1024>>>(g, d_out); break;
            case 1: rate_kernel<1><<<1, 128, 200 * 1024>>>(g, d_out); break;
            case 2: rate_kernel<2><<<1, 128, 200 * 1024>>>(g, d_out); break;
            case 3: rate_kernel<3><<<1, 128, 200 * 1024>>>(g, d_out); break;
            default: rate_kernel<6><<<1, 128, 200 * 1024>>>(g, d_out); break;
          }
          CK(cudaDeviceSynchronize());
          long long h[2];
          CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
          const int commits = pc ? g.n_instr / pc : iss;
          printf("commit cost: %s %d, %2d MMAs per commit: %u MMAs + %d commits in %6lld cycles (%.1f per op, %.1f per MMA)\n", lanes_mode ? "lanes" : "warps", iss, pc,
                 g.n_instr, commits, h[1], (double)h[1] / (g.n_instr + commits), (double)h[1] / g.n_instr);
        }
  }
  return 0;
}
