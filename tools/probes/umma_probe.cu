// umma_probe.cu — stand-alone probe of tcgen05.mma kind::i8 on sm_100a: which shared-memory operand layouts /
// descriptors give the product we expect, and where the accumulator rows land in tensor memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu ; run on a B200.
// The host lays each operand out byte by byte (so every layout hypothesis is explicit), the kernel only copies
// the images to shared memory, issues the MMAs and dumps all 128 lanes of the accumulator columns.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Probe {
  uint32_t a_bytes, b_bytes;        // operand image sizes
  uint64_t a_desc, b_desc;          // descriptors with start address 0
  uint32_t a_step, b_step;          // bytes the start address advances per K step
  uint32_t ksteps, idesc, ncols;    // ncols: accumulator columns to dump (multiple of 8)
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, Probe p, int32_t* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  uint8_t* g = smem + (sbase - (uint32_t)__cvta_generic_to_shared(smem));
  const uint32_t a_off = 0, b_off = (p.a_bytes + 1023u) & ~1023u;
  for (uint32_t i = threadIdx.x; i < p.a_bytes; i += 128) g[a_off + i] = a_img[i];
  for (uint32_t i = threadIdx.x; i < p.b_bytes; i += 128) g[b_off + i] = b_img[i];
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr) : "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (threadIdx.x == 0) {
    for (uint32_t k = 0; k < p.ksteps; k++) {
      const uint64_t ad = p.a_desc + (uint64_t)(((sbase + a_off + k * p.a_step) >> 4) & 0x3FFF);
      const uint64_t bd = p.b_desc + (uint64_t)(((sbase + b_off + k * p.b_step) >> 4) & 0x3FFF);
      const uint32_t acc = k > 0;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tm), "l"(ad), "l"(bd), "r"(p.idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
  }
  asm volatile(
      "{\n.reg .pred q;\nW: mbarrier.try_wait.parity.shared::cta.b64 q, [%0], 0;\n@q bra D;\nbra W;\nD:\n}\n" ::"r"(bar_addr)
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t warp = threadIdx.x >> 5;
  for (uint32_t c = 0; c < p.ncols; c += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tm + ((warp * 32u) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; j++) out[threadIdx.x * p.ncols + c + j] = (int32_t)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

static uint32_t make_idesc(int M, int N, int a_signed, int b_signed, int a_mn, int b_mn) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
static uint64_t make_sdesc(uint32_t lbo, uint32_t sbo, int layout) {
  return ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

// operand element -> byte offset, for the layouts under test
struct Lay { int kind; int rows, K; };   // rows = M or N extent
// kind 0: K-major, no swizzle:   [k/16][r/8][r%8][16]            LBO = rows*16, SBO = 128, step(32 K) = 2*LBO
// kind 1: MN-major, no swizzle:  [r/16][k/8][k%8][16 (r%16)]     LBO = 128 (next 8 k), SBO = K*16 (next 16 r), step = 4*128
// kind 2: MN-major, 128B swizzle:[r/128][k][128 (r%128)] with 16-byte chunk ^= k%8;  LBO = K*128 (next 128 r), SBO = 1024, step = 32*128
// kind 3: K-major, 128B swizzle: [r][128 (k%128)] chunk ^= r%8 (K <= 128);  LBO = 16 (unused), SBO = 1024, step = 32 bytes
static size_t lay_off(const Lay& L, int r, int k) {
  switch (L.kind) {
    case 0: return (size_t)(k / 16) * L.rows * 16 + (size_t)(r / 8) * 128 + (r % 8) * 16 + k % 16;
    case 1: return (size_t)(r / 16) * L.K * 16 + (size_t)(k / 8) * 128 + (k % 8) * 16 + r % 16;
    case 2: return (size_t)(r / 128) * L.K * 128 + (size_t)k * 128 + ((((r % 128) / 16) ^ (k % 8)) * 16) + r % 16;
    default: return (size_t)r * 128 + (((k / 16) ^ (r % 8)) * 16) + k % 16;
  }
}
static size_t lay_bytes(const Lay& L) { return L.kind == 3 ? (size_t)L.rows * 128 : (size_t)((L.rows + 127) / 128 * 128) * L.K; }
static void lay_desc(const Lay& L, uint64_t* desc, uint32_t* step) {
  switch (L.kind) {
    case 0: *desc = make_sdesc(L.rows * 16, 128, 0); *step = 2 * L.rows * 16; break;
    case 1: *desc = make_sdesc(128, L.K * 16, 0); *step = 4 * 128; break;
    case 2: *desc = make_sdesc(L.K * 128, 1024, 2); *step = 32 * 128; break;
    default: *desc = make_sdesc(16, 1024, 2); *step = 32; break;
  }
}

int main() {
  int dev_smem = 200 * 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dev_smem));
  struct Case { const char* name; int M, N, K, akind, bkind; };
  const Case cases[] = {
      {"A Kmaj none  x B Kmaj none, M128", 128, 64, 64, 0, 0},
      {"A Kmaj none  x B Kmaj none, M64 ", 64, 64, 64, 0, 0},
      {"A MNmaj none x B Kmaj none, M128", 128, 64, 64, 1, 0},
      {"A MNmaj sw128 x B Kmaj none, M128", 128, 64, 64, 2, 0},
      {"A Kmaj none  x B MNmaj none, M128", 128, 64, 64, 0, 1},
      {"A Kmaj none  x B MNmaj sw128, M128 N128", 128, 128, 64, 0, 2},
      {"A Kmaj sw128 x B Kmaj sw128, M128", 128, 64, 64, 3, 3},
      {"A MNmaj sw128 x B Kmaj none, M128 N256 K160", 128, 256, 160, 2, 0},
      {"A MNmaj none x B Kmaj none, M64 N224 K256", 64, 224, 256, 1, 0},
  };
  uint8_t *d_a, *d_b;
  int32_t* d_out;
  CK(cudaMalloc(&d_a, 1 << 20));
  CK(cudaMalloc(&d_b, 1 << 20));
  CK(cudaMalloc(&d_out, 128 * 512 * 4));
  for (const Case& c : cases) {
    Lay la{c.akind, c.M, c.K}, lb{c.bkind, c.N, c.K};
    std::vector<uint8_t> ai(lay_bytes(la), 0), bi(lay_bytes(lb), 0);
    std::vector<int> A((size_t)c.M * c.K), B((size_t)c.N * c.K);
    srand(1234);
    for (int m = 0; m < c.M; m++)
      for (int k = 0; k < c.K; k++) {
        A[(size_t)m * c.K + k] = rand() % 256;                       // u8
        ai[lay_off(la, m, k)] = (uint8_t)A[(size_t)m * c.K + k];
      }
    for (int n = 0; n < c.N; n++)
      for (int k = 0; k < c.K; k++) {
        B[(size_t)n * c.K + k] = rand() % 256 - 128;                 // s8
        bi[lay_off(lb, n, k)] = (uint8_t)(int8_t)B[(size_t)n * c.K + k];
      }
    Probe p;
    p.a_bytes = (uint32_t)ai.size();
    p.b_bytes = (uint32_t)bi.size();
    lay_desc(la, &p.a_desc, &p.a_step);
    lay_desc(lb, &p.b_desc, &p.b_step);
    p.ksteps = c.K / 32;
    p.idesc = make_idesc(c.M, c.N, 0, 1, c.akind == 1 || c.akind == 2, c.bkind == 1 || c.bkind == 2);
    p.ncols = c.N;
    CK(cudaMemcpy(d_a, ai.data(), ai.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, bi.data(), bi.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xEE, 128 * 512 * 4));
    probe_kernel<<<1, 128, dev_smem>>>(d_a, d_b, p, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-46s: launch failed: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    std::vector<int32_t> out((size_t)128 * c.N);
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    // expected D[m][n]; find which lane holds row m (try identity, and report the best lane permutation guess)
    std::vector<long long> D((size_t)c.M * c.N);
    for (int m = 0; m < c.M; m++)
      for (int n = 0; n < c.N; n++) {
        long long s = 0;
        for (int k = 0; k < c.K; k++) s += (long long)A[(size_t)m * c.K + k] * B[(size_t)n * c.K + k];
        D[(size_t)m * c.N + n] = s;
      }
    int ok_rows = 0;
    std::vector<int> lane_of(c.M, -1);
    for (int m = 0; m < c.M; m++)
      for (int lane = 0; lane < 128; lane++) {
        bool eq = true;
        for (int n = 0; n < c.N && eq; n++) eq = out[(size_t)lane * c.N + n] == D[(size_t)m * c.N + n];
        if (eq) { lane_of[m] = lane; ok_rows++; break; }
      }
    printf("%-46s: %d / %d rows found;", c.name, ok_rows, c.M);
    if (ok_rows == c.M) {
      bool ident = true;
      for (int m = 0; m < c.M; m++) ident &= lane_of[m] == m;
      if (ident) printf(" row m -> lane m\n");
      else { printf(" lanes:"); for (int m = 0; m < c.M; m += c.M / 8) printf(" %d->%d", m, lane_of[m]); printf("\n"); }
    } else {
      printf(" first lanes: out[0][0..3] = %d %d %d %d expected %lld %lld %lld %lld\n", out[0], out[1], out[2], out[3], D[0], D[1], D[2], D[3]);
    }
  }
  return 0;
}
