// Microbenchmark (bring-up tool, not product): TMEM read (tcgen05.ld) throughput alone and while the
// tensor core is running tcgen05.mma kind::i8 into another TMEM region.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../scanner_colmap_b200/csrc/ptx.cuh"
using namespace smb::ptx;

__device__ __forceinline__ void ld_x32(uint32_t taddr) {
  uint32_t v[32];
  tmem_ld_32x32b_x32(taddr, v);
}
__device__ __forceinline__ void ld_x64(uint32_t taddr) {
  asm volatile(
      "{ .reg .b32 r<64>;\n"
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{r0,r1,r2,r3,r4,r5,r6,r7,r8,r9,r10,r11,r12,r13,r14,r15,r16,r17,r18,r19,r20,r21,r22,r23,r24,r25,r26,r27,r28,r29,r30,r31,"
      "r32,r33,r34,r35,r36,r37,r38,r39,r40,r41,r42,r43,r44,r45,r46,r47,r48,r49,r50,r51,r52,r53,r54,r55,r56,r57,r58,r59,r60,r61,r62,r63}, [%0]; }"
      ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_16x256b_x8(uint32_t taddr) {  // 16 lanes x 256 bit, repeated 8x -> 32 regs
  asm volatile(
      "{ .reg .b32 r<32>;\n"
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{r0,r1,r2,r3,r4,r5,r6,r7,r8,r9,r10,r11,r12,r13,r14,r15,r16,r17,r18,r19,r20,r21,r22,r23,r24,r25,r26,r27,r28,r29,r30,r31}, [%0]; }"
      ::"r"(taddr) : "memory");
}

// mode: 0 = x32, 1 = x64, 2 = 16x256b.x8
__global__ void __launch_bounds__(640, 1) mix_kernel(int mma_iters, int ld_iters, int n_ld_warps, int mode, int waits_every,
                                                      long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (uint32_t x = threadIdx.x; x < (16384 + 32768) / 4; x += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (smem0 - smem_u32(smem_raw)))[x] = 0x01010101u;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 2) { tmem_alloc_512(smem_u32(&tmem_base_s)); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = tmem_base_s;
  if (warp == 1 && lane == 0 && mma_iters > 0) {
    const uint64_t adesc = make_kmajor_sw128_desc(smem0), bdesc = make_kmajor_sw128_desc(smem0 + 16384);
    const uint32_t idesc = make_idesc_u8u8s32(128, 256);
    long long t0 = clock64();
    for (int i = 0; i < mma_iters; ++i)
      for (int k = 0; k < 4; ++k) umma_i8(tb, adesc + k * 2, bdesc + k * 2, idesc, k);   // always columns 0..255
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    out[blockIdx.x * 16 + 0] = clock64() - t0;
  }
  if (warp >= 4 && warp < 4 + n_ld_warps && ld_iters > 0) {
    const uint32_t lane_addr = ((warp & 3) * 32u) << 16;
    const uint32_t colbase = 256 + (((warp - 4) >> 2) & 1) * 128;    // columns 256..511: the "other" accumulator
    long long t0 = clock64();
    for (int i = 0; i < ld_iters; ++i) {
      const uint32_t ta = tb + lane_addr + colbase + (i & 1) * 64;
      if (mode == 0) { ld_x32(ta); ld_x32(ta + 32); }
      else if (mode == 1) ld_x64(ta);
      else { ld_16x256b_x8(ta); ld_16x256b_x8(ta + 32); }   // NB: different lane coverage, bandwidth probe only
      if ((i + 1) % waits_every == 0) tmem_wait_ld();
    }
    tmem_wait_ld();
    if (warp < 16) out[blockIdx.x * 16 + warp] = clock64() - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) { tcgen05_fence_after(); tmem_dealloc_512(tb); }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 1024 + 16384 + 32768;
  cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d; cudaMalloc(&d, sms * 16 * sizeof(long long));
  long long h[148 * 16 + 64];
  const int mma_iters = 1000, ld_iters = 4000;
  struct Cfg { int mma, ld, warps, mode, waits; const char* name; } cfgs[] = {
    {0, 1, 4, 0, 1, "ld x32 x2, 4 warps, wait each"}, {0, 1, 8, 0, 1, "ld x32 x2, 8 warps, wait each"},
    {0, 1, 12, 0, 1, "ld x32 x2, 12 warps, wait each"}, {0, 1, 16, 0, 1, "ld x32 x2, 16 warps, wait each"},
    {1, 1, 16, 0, 1, "mma + ld x32x2 16 warps"}, {1, 1, 12, 0, 1, "mma + ld x32x2 12 warps"},
    {0, 1, 4, 0, 4, "ld x32 x2, 4 warps, wait/4"},   {0, 1, 8, 0, 4, "ld x32 x2, 8 warps, wait/4"},
    {0, 1, 4, 1, 1, "ld x64, 4 warps, wait each"},   {0, 1, 8, 1, 1, "ld x64, 8 warps, wait each"},
    {0, 1, 8, 2, 1, "ld 16x256b.x8 x2, 8 warps"},
    {1, 0, 0, 0, 1, "mma only"},
    {1, 1, 4, 0, 1, "mma + ld x32x2 4 warps"}, {1, 1, 8, 0, 1, "mma + ld x32x2 8 warps"}, {1, 1, 8, 1, 1, "mma + ld x64 8 warps"},
    {1, 1, 8, 2, 1, "mma + ld 16x256b 8 warps"}, {1, 1, 2, 0, 1, "mma + ld x32x2 2 warps"}, {1, 1, 1, 0, 1, "mma + ld x32x2 1 warp"},
  };
  for (auto& c : cfgs) {
    cudaMemset(d, 0, sms * 16 * sizeof(long long));
    // size the ld loop so both roles run for a similar time when mixed
    for (int rep = 0; rep < 2; ++rep)
      mix_kernel<<<sms, 640, smem>>>(c.mma ? mma_iters : 0, c.ld ? ld_iters : 0, c.warps, c.mode, c.waits, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: ERROR %s\n", c.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sms * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mma_c = 0, ld_c = 0;
    for (int b = 0; b < sms; ++b) { mma_c = h[b*16] > mma_c ? h[b*16] : mma_c; for (int w = 4; w < 16; ++w) ld_c = h[b*16+w] > ld_c ? h[b*16+w] : ld_c; }
    printf("%-34s", c.name);
    if (c.mma) printf(" mma: %7.1f cyc/MMA (128 ideal)", (double)mma_c / (mma_iters * 4.0));
    if (c.ld) printf("  ld: %7.1f cyc per 64 cols per warp -> %6.1f B/clk/SM", (double)ld_c / ld_iters,
                     (double)c.warps * 64 * 32 * 4 / ((double)ld_c / ld_iters));
    printf("\n");
  }
  return 0;
}
