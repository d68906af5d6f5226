// Microbenchmark (bring-up tool, not product): issue rate of tcgen05.mma for several kinds / shapes on
// sm_100a, one CTA per SM, one issuing thread, operands in (uninitialised) swizzled smem.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate tools/umma_rate.cu && /tmp/umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../scanner_colmap_b200/csrc/ptx.cuh"

using namespace smb::ptx;

enum Kind { I8 = 0, F8 = 1, F16 = 2 };

template <int KIND>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == I8)
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p; }" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else if (KIND == F8)
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p; }" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

__device__ int g_fill_mode = 0;
__device__ int g_commit_every = 0;  // 0: one commit at the end; n: commit to a scratch barrier after every n-th group of k_steps MMAs  // 0: constant bytes, 1: pseudo-random bytes, 2: zeros
template <int KIND>
__global__ void __launch_bounds__(128, 1) rate_kernel(uint32_t idesc, int iters, int n_cols, int k_steps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // zero the operand tiles so fp kinds see no NaN patterns
  for (uint32_t x = threadIdx.x; x < (16384 + 32768) / 4; x += 128)
    reinterpret_cast<uint32_t*>(smem_raw + (smem0 - smem_u32(smem_raw)))[x] =
        g_fill_mode == 0 ? 0x01010101u : g_fill_mode == 2 ? 0u : ((x * 2654435761u) ^ (x >> 3) * 40503u) & (KIND == I8 ? 0xFFFFFFFFu : 0x3F3F3F3Fu);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar2), 1u << 20);  // never completes: a pure commit target
    fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc_512(smem_u32(&tmem_base_s));
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint64_t adesc = make_kmajor_sw128_desc(smem0);
    const uint64_t bdesc = make_kmajor_sw128_desc(smem0 + 16384);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tb + (n_cols <= 160 ? (i % 3) : (i & 1)) * n_cols;
      for (int k = 0; k < k_steps; ++k) mma<KIND>(d, adesc + k * 2, bdesc + k * 2, idesc, k);
      if (g_commit_every && (i % g_commit_every) == g_commit_every - 1) umma_commit(smem_u32(&bar2));
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tcgen05_fence_after();
    tmem_dealloc_512(tb);
  }
}

static uint32_t idesc_of(int kind, int afmt, int bfmt, int m, int n) {
  uint32_t cfmt = kind == I8 ? 2u : 1u;
  return (cfmt << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int KIND>
void run(const char* name, int afmt, int bfmt, int m, int n, int k_steps, int grid) {
  long long* d;
  cudaMalloc(&d, grid * sizeof(long long));
  const int smem = 1024 + 16384 + 32768;
  cudaFuncSetAttribute(rate_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  const uint32_t idesc = idesc_of(KIND, afmt, bfmt, m, n);
  for (int rep = 0; rep < 2; ++rep) rate_kernel<KIND><<<grid, 128, smem>>>(idesc, iters, n, k_steps, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-28s ERROR %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  long long h[256];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double cyc_per_mma = (double)mx / ((double)iters * k_steps);
  const double kbytes = 32.0;  // bytes of K per instruction for every kind here
  const double k_elems = KIND == F16 ? 16.0 : kbytes;
  const double macs = (double)m * n * k_elems;
  printf("%-28s grid=%3d M=%d N=%3d ksteps=%d: %7.1f cycles/MMA  -> %6.0f MAC/clk/SM\n", name, grid, m, n, k_steps, cyc_per_mma,
         macs / cyc_per_mma);
  cudaFree(d);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int ce : {0, 1, 2})
  for (int grid : {sms}) {
    int mode = 1;
    cudaMemcpyToSymbol(g_fill_mode, &mode, sizeof(int));
    cudaMemcpyToSymbol(g_commit_every, &ce, sizeof(int));
    printf("--- commit every %d group(s) of 4 MMAs (0 = only at the end)\n", ce);
    run<I8>("i8  u8 x u8", 0, 0, 128, 256, 4, grid);
    run<I8>("i8  s8 x s8", 1, 1, 128, 256, 4, grid);
    run<I8>("i8  u8 x u8 N=128", 0, 0, 128, 128, 4, grid);
    run<I8>("i8  u8 x u8 N=160", 0, 0, 128, 160, 4, grid);
    run<I8>("i8  u8 x u8 N=192", 0, 0, 128, 192, 4, grid);
    run<I8>("i8  u8 x u8 N=224", 0, 0, 128, 224, 4, grid);
    run<I8>("i8  u8 x u8 M=64", 0, 0, 64, 256, 4, grid);
    run<F8>("f8f6f4 e4m3", 0, 0, 128, 256, 4, grid);
    run<F16>("f16 (K=16)", 0, 0, 128, 256, 4, grid);
    run<F16>("bf16 (K=16)", 1, 1, 128, 256, 4, grid);
  }
  return 0;
}
