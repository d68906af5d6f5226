// Microbenchmark (bring-up tool, not product): does the score kernel's OTHER shared-memory / TMEM traffic slow the
// tensor pipe down?  One CTA per SM, the product's role layout, but no hand-offs at all: the MMA thread issues
// tiles back to back (nothing waits for anything), and the other roles generate their traffic at a clock-paced
// rate.  Reported: clk per 128 x N accumulator tile as seen by the MMA thread (512 = tensor pipe time at N = 256).
//   roles   warp 1 lane 0 : tcgen05.mma kind::i8, 4 K-steps per tile, A from smem (SS) or from TMEM (TS),
//                           B from a 4-stage ring, two tiles (the two M halves of a strip) per B stage
//           warp 0 lane 0 : cp.async.bulk global -> B ring, `tma_bytes` per B stage every `tma_pace` clk
//           warps 4..11   : tcgen05.ld 32x32b.x32 x 4 per tile every `ld_pace` clk (the epilogue's TMEM read)
//           warps 2..3    : st.shared.v4 + ld.shared.v4 of `mail_bytes` per tile every `ld_pace` clk (mailbox traffic)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/smem_port tools/smem_port.cu && /tmp/smem_port
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../scanner_colmap_b200/csrc/ptx.cuh"

using namespace smb::ptx;

struct Params {
  int iters;       // B stages (two tiles each)
  int n_cols;      // MMA N
  int a_tmem;      // 1: A operand from TMEM (columns 448..511)
  int tma_bytes;   // per B stage, 0 = no TMA traffic
  int tma_pace;    // clk per B stage (0 = as fast as the ring allows)
  int ldtm;        // 1: epilogue warps read TMEM
  int ld_pace;     // clk per tile for the epilogue / mailbox roles
  int mail_bytes;  // per tile, written and read back by warps 2..3
  int probe;       // 1: warps 4..11 time dependent shared-memory operations instead of reading TMEM
};

__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p; }" ::"r"(tmem_d),
               "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
               : "memory");
}

__device__ __forceinline__ void pace_until(long long t) {
  while (clock64() < t) {
  }
}

constexpr int kStageBytes = 32768, kStages = 4, kABytes = 32768, kMailBytes = 16384;

__global__ void __launch_bounds__(384, 1) port_kernel(Params p, const uint8_t* gsrc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar_done, bar_full[kStages], bar_ready;
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (smem0 - smem_u32(smem_raw));
  for (uint32_t x = threadIdx.x; x < (kABytes + kStages * kStageBytes + kMailBytes) / 4; x += blockDim.x)
    reinterpret_cast<uint32_t*>(base)[x] = (x * 2654435761u) ^ ((x >> 3) * 40503u);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar_done), 1);
    mbar_init(smem_u32(&bar_ready), 1);
    mbar_arrive(smem_u32(&bar_ready));  // phase 0 complete: a wait on parity 0 succeeds at once
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&bar_full[s]), 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_512(smem_u32(&tmem_base_s));
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = tmem_base_s;
  const long long t_start = clock64();
  if (warp == 1 && lane == 0 && p.a_tmem != 2) {
    const uint64_t adesc2[2] = {make_kmajor_sw128_desc(smem0), make_kmajor_sw128_desc(smem0 + 16384)};
    const uint32_t idesc = make_idesc_u8u8s32(128, (uint32_t)p.n_cols);
    const long long t0 = clock64();
    for (int i = 0; i < p.iters; ++i) {
      const uint64_t bdesc = make_kmajor_sw128_desc(smem0 + kABytes + (i % kStages) * kStageBytes);
      for (int m = 0; m < 2; ++m) {
        const uint32_t d = tb + m * p.n_cols;
        if (p.a_tmem)
          for (int k = 0; k < 4; ++k) umma_i8_ts(d, tb + 448 + m * 32 + k * 8, bdesc + k * 2, idesc, k);
        else
          for (int k = 0; k < 4; ++k) umma_i8(d, adesc2[m] + k * 2, bdesc + k * 2, idesc, k);
      }
    }
    umma_commit(smem_u32(&bar_done));
    mbar_wait(smem_u32(&bar_done), 0);
    out[blockIdx.x * 16 + 0] = clock64() - t0;
  }
  if (warp == 0 && lane == 0 && p.tma_bytes > 0) {
    const long long t0 = clock64();
    for (int i = 0; i < p.iters; ++i) {
      const int s = i % kStages;
      if (i >= kStages) mbar_wait(smem_u32(&bar_full[s]), ((i / kStages) - 1) & 1);  // the stage's previous load has landed
      if (p.tma_pace) pace_until(t_start + (long long)i * p.tma_pace);
      mbar_arrive_expect_tx(smem_u32(&bar_full[s]), (uint32_t)p.tma_bytes);
      const uint8_t* src = gsrc + ((size_t)i * kStageBytes) % (1u << 20);
      for (int off = 0; off < p.tma_bytes; off += 16384)
        bulk_load(smem0 + kABytes + s * kStageBytes + off, src + off, (uint32_t)min(16384, p.tma_bytes - off), smem_u32(&bar_full[s]));
    }
    for (int i = max(0, p.iters - kStages); i < p.iters; ++i) mbar_wait(smem_u32(&bar_full[i % kStages]), (i / kStages) & 1);
    out[blockIdx.x * 16 + 1] = clock64() - t0;
  }
  if (warp >= 4 && p.ldtm) {
    const uint32_t ta = tb + (((warp & 3) * 32u) << 16) + ((warp - 4) >> 2) * 128;
    uint32_t sink = 0;
    for (int i = 0; i < 2 * p.iters; ++i) {
      pace_until(t_start + (long long)i * p.ld_pace);
      uint32_t v0[32], v1[32], v2[32], v3[32];
      const uint32_t t = ta + (i & 1) * p.n_cols;  // alternate between the two accumulators, as the epilogue does
      tmem_ld_32x32b_x32(t, v0);
      tmem_ld_32x32b_x32(t + 32, v1);
      tmem_ld_32x32b_x32(t + 64, v2);
      tmem_ld_32x32b_x32(t + 96, v3);
      tmem_wait_ld();
#pragma unroll
      for (int x = 0; x < 32; ++x) sink = max(sink, max(max(v0[x], v1[x]), max(v2[x], v3[x])));
    }
    if (sink == 0x12345678u) out[blockIdx.x * 16 + 3] = sink;
    if (lane == 0 && warp == 4) out[blockIdx.x * 16 + 2] = clock64() - t_start;
  }
  if (warp >= 4 && p.probe) {
    // latencies an epilogue warp sees while the tensor core streams its operands: (a) try_wait on a complete
    // mbarrier, (b) STS.128 -> LDS.128 round trip, (c) a dependent SHFL chain, (d) mbarrier arrive (+ dependent wait)
    uint4* mb = reinterpret_cast<uint4*>(base + kABytes + kStages * kStageBytes) + (warp - 4) * 32;
    const int reps = 1000;
    pace_until(t_start + 20000);  // let the MMA stream reach steady state
    long long t0 = clock64();
    uint32_t ok = 0;
    for (int i = 0; i < reps; ++i) ok += mbar_try_wait(smem_u32(&bar_ready), 0);
    long long t1 = clock64();
    uint4 v = make_uint4(lane, 1, 2, 3);
    for (int i = 0; i < reps; ++i) {
      mb[lane] = v;
      __syncwarp();
      const uint4 w = mb[(lane + 1) & 31];
      v.x += w.y + (ok & 1);
      __syncwarp();
    }
    long long t2 = clock64();
    uint32_t sh = v.x;
    for (int i = 0; i < reps; ++i) sh = __shfl_xor_sync(0xffffffffu, sh, 1) + i;
    long long t3 = clock64();
    if (lane == 0 && warp == 4 && sh != 0x12345678u) {
      out[blockIdx.x * 16 + 4] = t1 - t0;
      out[blockIdx.x * 16 + 5] = t2 - t1;
      out[blockIdx.x * 16 + 6] = t3 - t2;
    }
  }
  if ((warp == 2 || warp == 3) && p.mail_bytes > 0) {
    // two warps: each moves mail_bytes / 2 per tile (16 B per lane and instruction = 512 B per warp instruction)
    const int n_inst = p.mail_bytes / 2 / 512;
    uint4* mb = reinterpret_cast<uint4*>(base + kABytes + kStages * kStageBytes) + (warp - 2) * 512;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < 2 * p.iters; ++i) {
      pace_until(t_start + (long long)i * p.ld_pace);
      for (int x = 0; x < n_inst; ++x) mb[(x & 15) * 32 + lane] = make_uint4(i, x, lane, acc.x);
      __syncwarp();
      for (int x = 0; x < n_inst; ++x) {
        const uint4 v = mb[(x & 15) * 32 + ((lane + 1) & 31)];
        acc.x ^= v.x + v.y;
      }
    }
    if (acc.x == 0x12345678u) out[blockIdx.x * 16 + 3] = acc.x;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_512(tb);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 1024 + kABytes + kStages * kStageBytes + kMailBytes;
  cudaFuncSetAttribute(port_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  cudaMalloc(&d, sms * 16 * sizeof(long long));
  uint8_t* g;
  cudaMalloc(&g, 2u << 20);
  cudaMemset(g, 1, 2u << 20);
  const int iters = 1000;
  struct Cfg {
    const char* name;
    Params p;
  } cfgs[] = {
      //                                  iters  N  aT  tmaB   pace ldtm ldp  mail
      {"SS N=256 alone", {iters, 256, 0, 0, 0, 0, 512, 0, 0}},
      {"SS N=256 + TMA 32K/1024clk", {iters, 256, 0, 32768, 1024, 0, 512, 0, 0}},
      {"SS N=256 + TMA 32K/512clk (2x rate)", {iters, 256, 0, 32768, 512, 0, 512, 0, 0}},
      {"SS N=256 + TMA unpaced", {iters, 256, 0, 32768, 0, 0, 512, 0, 0}},
      {"SS N=256 + LDTM 8 warps/512clk", {iters, 256, 0, 0, 0, 1, 512, 0, 0}},
      {"SS N=256 + mailbox 16K/tile", {iters, 256, 0, 0, 0, 0, 512, 16384, 0}},
      {"SS N=256 + TMA + LDTM", {iters, 256, 0, 32768, 1024, 1, 512, 0, 0}},
      {"SS N=256 + TMA + LDTM + mailbox 4K", {iters, 256, 0, 32768, 1024, 1, 512, 4096, 0}},
      {"SS N=256 + TMA + LDTM + mailbox 16K", {iters, 256, 0, 32768, 1024, 1, 512, 16384, 0}},
      {"SS N=224 alone", {iters, 224, 0, 0, 0, 0, 448, 0, 0}},
      {"SS N=224 + TMA 28K/896clk", {iters, 224, 0, 28672, 896, 0, 448, 0, 0}},
      {"SS N=224 + TMA + LDTM + mailbox 4K", {iters, 224, 0, 28672, 896, 1, 448, 4096, 0}},
      {"TS N=224 alone", {iters, 224, 1, 0, 0, 0, 448, 0, 0}},
      {"TS N=224 + TMA 28K/896clk", {iters, 224, 1, 28672, 896, 0, 448, 0, 0}},
      {"TS N=224 + TMA + LDTM", {iters, 224, 1, 28672, 896, 1, 448, 0, 0}},
      {"TS N=224 + TMA + LDTM + mailbox 4K", {iters, 224, 1, 28672, 896, 1, 448, 4096, 0}},
      {"TS N=224 + TMA + LDTM + mailbox 16K", {iters, 224, 1, 28672, 896, 1, 448, 16384, 0}},
      {"TS N=224 + TMA unpaced", {iters, 224, 1, 28672, 0, 0, 448, 0, 0}},
      {"probe: no MMA", {1, 256, 0, 0, 0, 0, 512, 0, 1}},
      {"probe: no MMA + TMA 32K/1024", {1000, 256, 2, 32768, 1024, 0, 512, 0, 1}},
      {"probe: SS N=256 + TMA", {iters, 256, 0, 32768, 1024, 0, 512, 0, 1}},
      {"probe: SS N=224 + TMA", {iters, 224, 0, 28672, 896, 0, 448, 0, 1}},
      {"probe: TS N=224 + TMA", {iters, 224, 1, 28672, 896, 0, 448, 0, 1}},
      {"probe: TS N=224 + TMA + mailbox 4K", {iters, 224, 1, 28672, 896, 0, 448, 4096, 1}},
      {"probe: SS N=256 + TMA + mailbox 4K", {iters, 256, 0, 32768, 1024, 0, 512, 4096, 1}},
      {"SS N=128 alone", {iters, 128, 0, 0, 0, 0, 256, 0, 0}},
      {"TS N=128 alone", {iters, 128, 1, 0, 0, 0, 256, 0, 0}},
      {"TS N=128 + TMA 16K/512clk + LDTM", {iters, 128, 1, 16384, 512, 1, 256, 0, 0}},
  };
  long long h[148 * 16 + 16];
  for (auto& c : cfgs) {
    cudaMemset(d, 0, sms * 16 * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) port_kernel<<<sms, 384, smem>>>(c.p, g, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: ERROR %s\n", c.name, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(h, d, sms * 16 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0, tma = 0;
    double mean = 0;
    for (int b = 0; b < sms; ++b) {
      mx = h[b * 16] > mx ? h[b * 16] : mx;
      mean += (double)h[b * 16] / sms;
      tma = h[b * 16 + 1] > tma ? h[b * 16 + 1] : tma;
    }
    const double tiles = 2.0 * c.p.iters;
    printf("%-40s mma: %6.1f clk/tile max, %6.1f mean (ideal %d)", c.name, mx / tiles, mean / tiles, 2 * c.p.n_cols);
    if (c.p.probe) {
      double a = 0, b2 = 0, c2 = 0;
      for (int b = 0; b < sms; ++b) { a += (double)h[b * 16 + 4] / sms; b2 += (double)h[b * 16 + 5] / sms; c2 += (double)h[b * 16 + 6] / sms; }
      printf("   probe (clk, mean over CTAs): try_wait %5.1f  sts+lds %5.1f  shfl %5.1f", a / 1000, b2 / 1000, c2 / 1000);
    }
    if (c.p.tma_bytes) printf("   tma: %6.1f clk per stage -> %5.1f B/clk", (double)tma / c.p.iters, (double)c.p.tma_bytes * c.p.iters / tma);
    printf("\n");
  }
  return 0;
}
