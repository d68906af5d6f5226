// Microbenchmark (bring-up tool): latency of tcgen05.commit -> mbarrier -> waiting thread, on sm_100a.
//  mode 0: [4 MMAs, commit, wait] fully serialized by the issuing thread         -> 4*128 + commit/wake latency
//  mode 1: same, but another warp waits on the barrier and then arrives on a second barrier the issuer
//          waits on (the MMA -> epilogue -> MMA hand-off with an epilogue that does nothing)
//  mode 2: two buffers ping-pong through mode 1's hand-off (what the score kernel's pipeline does)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../scanner_colmap_b200/csrc/ptx.cuh"
using namespace smb::ptx;

__global__ void __launch_bounds__(128, 1) k(int mode, int iters, int nwaiters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full[2], empty[2];
  __shared__ uint32_t tb_s;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (uint32_t x = threadIdx.x; x < (16384 + 32768) / 4; x += 128)
    reinterpret_cast<uint32_t*>(smem_raw + (smem0 - smem_u32(smem_raw)))[x] = x * 2654435761u;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&full[b]), 1); mbar_init(smem_u32(&empty[b]), nwaiters); }
    fence_barrier_init();
  }
  if (warp == 3) { tmem_alloc_512(smem_u32(&tb_s)); tmem_relinquish(); }
  fence_proxy_async_smem();
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tb = tb_s;
  const uint64_t adesc = make_kmajor_sw128_desc(smem0), bdesc = make_kmajor_sw128_desc(smem0 + 16384);
  const uint32_t idesc = make_idesc_u8u8s32(128, 256);
  if (warp == 0 && lane == 0) {
    long long t0 = clock64();
    if (mode == 0) {
      for (int i = 0; i < iters; ++i) {
        for (int kk = 0; kk < 4; ++kk) umma_i8(tb, adesc + kk * 2, bdesc + kk * 2, idesc, kk);
        umma_commit(smem_u32(&full[0]));
        mbar_wait(smem_u32(&full[0]), i & 1);
      }
    } else if (mode == 1) {
      for (int i = 0; i < iters; ++i) {
        for (int kk = 0; kk < 4; ++kk) umma_i8(tb, adesc + kk * 2, bdesc + kk * 2, idesc, kk);
        umma_commit(smem_u32(&full[0]));
        mbar_wait(smem_u32(&empty[0]), i & 1);
        tcgen05_fence_after();
      }
    } else {
      uint32_t ts = 0, ph = 0;
      for (int i = 0; i < iters; ++i) {
        mbar_wait(smem_u32(&empty[ts]), ph ^ 1);
        tcgen05_fence_after();
        for (int kk = 0; kk < 4; ++kk) umma_i8(tb + ts * 256, adesc + kk * 2, bdesc + kk * 2, idesc, kk);
        umma_commit(smem_u32(&full[ts]));
        if (++ts == 2) { ts = 0; ph ^= 1; }
      }
      umma_commit(smem_u32(&full[ts]));  // drain
    }
    out[blockIdx.x] = clock64() - t0;
  } else if (warp >= 1 && warp <= (uint32_t)nwaiters && mode >= 1) {
    uint32_t ts = 0, ph = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(smem_u32(&full[mode == 1 ? 0 : ts]), mode == 1 ? (i & 1) : ph);
      tcgen05_fence_after();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[mode == 1 ? 0 : ts]));
      if (mode == 2 && ++ts == 2) { ts = 0; ph ^= 1; }
    }
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 3) { tcgen05_fence_after(); tmem_dealloc_512(tb); }
}

int main() {
  const int smem = 1024 + 16384 + 32768;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  long long h[148];
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int nw : {1, 2}) {
      if (mode == 0 && nw > 1) continue;
      k<<<148, 128, smem>>>(mode, iters, nw, d);
      k<<<148, 128, smem>>>(mode, iters, nw, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("mode %d waiters %d: %.1f cycles per tile of 4 MMAs (512 = tensor pipe time)\n", mode, nw, (double)mx / iters);
    }
  return 0;
}
