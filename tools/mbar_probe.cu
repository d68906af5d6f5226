// Probe (bring-up tool, not product): (1) the raw 64-bit contents of an mbarrier across phases, read with plain
// ld.shared, (2) DEPENDENT latencies of mbarrier.try_wait / test_wait on a complete phase vs a plain LDS of the
// barrier word, idle and while the tensor core runs (SS N=256 back to back).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../scanner_colmap_b200/csrc/ptx.cuh"
using namespace smb::ptx;

__device__ __forceinline__ unsigned long long lds64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.volatile.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(256, 1) probe(int with_mma, unsigned long long* raw, long long* lat) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar, bar_done, bar_c, bar_tx;
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t x = threadIdx.x; x < (16384 + 32768) / 4; x += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (smem0 - smem_u32(smem_raw)))[x] = x * 2654435761u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar_done), 1);
    mbar_init(smem_u32(&bar_c), 8);
    mbar_init(smem_u32(&bar_tx), 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc_512(smem_u32(&tmem_base_s)); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = tmem_base_s;
  if (threadIdx.x == 0 && blockIdx.x == 0 && !with_mma) {
    // raw contents: count-1 barrier through four phases; count-8 barrier through partial arrivals; tx-count pending
    int n = 0;
    raw[n++] = lds64(smem_u32(&bar));
    for (int i = 0; i < 4; ++i) { mbar_arrive(smem_u32(&bar)); raw[n++] = lds64(smem_u32(&bar)); }
    raw[n++] = lds64(smem_u32(&bar_c));
    for (int i = 0; i < 9; ++i) { mbar_arrive(smem_u32(&bar_c)); raw[n++] = lds64(smem_u32(&bar_c)); }
    raw[n++] = lds64(smem_u32(&bar_tx));
    mbar_arrive_expect_tx(smem_u32(&bar_tx), 4096);
    raw[n++] = lds64(smem_u32(&bar_tx));
    // commit-driven completion: tcgen05.commit with no MMA outstanding arrives promptly
    raw[n++] = lds64(smem_u32(&bar_done));
    umma_commit(smem_u32(&bar_done));
    mbar_wait(smem_u32(&bar_done), 0);
    raw[n++] = lds64(smem_u32(&bar_done));
  }
  __syncthreads();
  if (warp == 1 && lane == 0 && with_mma) {
    const uint64_t adesc = make_kmajor_sw128_desc(smem0), bdesc = make_kmajor_sw128_desc(smem0 + 16384);
    const uint32_t idesc = make_idesc_u8u8s32(128, 256);
    for (int i = 0; i < 1500; ++i)
      for (int k = 0; k < 4; ++k) umma_i8(tb + (i & 1) * 256, adesc + k * 2, bdesc + k * 2, idesc, k);
    umma_commit(smem_u32(&bar_done));
    mbar_wait(smem_u32(&bar_done), 0);
  }
  if (warp >= 4) {
    // `bar` has completed phase 0 (and more) in CTA 0's raw test; make every CTA's state the same: complete phase 0
    if (threadIdx.x == 128 && (blockIdx.x != 0 || with_mma)) mbar_arrive(smem_u32(&bar));
    __syncwarp();
    const uint32_t parity = (blockIdx.x == 0 && !with_mma) ? 1u : 0u;   // the most recently completed phase
    const uint32_t b = smem_u32(&bar);
    const int reps = 500;
    long long t0 = clock64();
    uint32_t dep = 0;
    for (int i = 0; i < reps; ++i) dep = mbar_try_wait(b + (dep & 8u), parity) - 1u + dep;   // dep stays 0 when it succeeds
    long long t1 = clock64();
    for (int i = 0; i < reps; ++i) dep = mbar_try_wait_hint(b + (dep & 8u), parity, 4000u) - 1u + dep;
    long long t2 = clock64();
    for (int i = 0; i < reps; ++i) dep = mbar_test(b + (dep & 8u), parity) - 1u + dep;
    long long t3 = clock64();
    for (int i = 0; i < reps; ++i) dep = (uint32_t)(lds64(b + (dep & 8u)) >> 63) * 0u + dep;
    long long t4 = clock64();
    unsigned long long acc = 0;
    for (int i = 0; i < reps; ++i) acc += lds64(b + (uint32_t)(acc & 8ull));
    long long t5 = clock64();
    if (lane == 0 && warp == 4) {
      lat[blockIdx.x * 8 + 0] = t1 - t0;
      lat[blockIdx.x * 8 + 1] = t2 - t1;
      lat[blockIdx.x * 8 + 2] = t3 - t2;
      lat[blockIdx.x * 8 + 3] = t4 - t3;
      lat[blockIdx.x * 8 + 4] = t5 - t4;
      lat[blockIdx.x * 8 + 5] = dep + (acc == 0x1234567 ? 1 : 0);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) { tcgen05_fence_after(); tmem_dealloc_512(tb); }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 1024 + 16384 + 32768;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  unsigned long long* raw; long long* lat;
  cudaMalloc(&raw, 64 * 8); cudaMalloc(&lat, sms * 8 * 8);
  cudaMemset(raw, 0, 64 * 8);
  for (int with_mma = 0; with_mma < 2; ++with_mma) {
    probe<<<sms, 256, smem>>>(with_mma, raw, lat);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148 * 8];
    cudaMemcpy(h, lat, sms * 8 * 8, cudaMemcpyDeviceToHost);
    double m[5] = {0, 0, 0, 0, 0};
    for (int b = 1; b < sms; ++b) for (int k = 0; k < 5; ++k) m[k] += (double)h[b * 8 + k] / (sms - 1) / 500.0;
    printf("%s: dependent latency (clk): try_wait %.1f  try_wait+hint %.1f  test_wait %.1f  lds64(dep on >>63) %.1f  lds64 chain %.1f   [dep=%lld]\n",
           with_mma ? "tensor core busy (SS N=256)" : "idle", m[0], m[1], m[2], m[3], m[4], h[8 + 5]);
    if (!with_mma) {
      unsigned long long r[32];
      cudaMemcpy(r, raw, 32 * 8, cudaMemcpyDeviceToHost);
      const char* names[] = {"count1 init", "count1 +1 arrive", "count1 +2", "count1 +3", "count1 +4",
                             "count8 init", "c8 +1", "c8 +2", "c8 +3", "c8 +4", "c8 +5", "c8 +6", "c8 +7", "c8 +8 (phase done)", "c8 +9",
                             "tx init", "tx after arrive.expect_tx 4096", "commit target init", "commit target after completion"};
      for (int i = 0; i < 19; ++i) printf("  raw %-34s 0x%016llx\n", names[i], r[i]);
    }
  }
  return 0;
}
