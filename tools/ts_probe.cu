// Functional probe (bring-up tool, not product): tcgen05.mma kind::i8 with the A operand in TENSOR MEMORY.
// One CTA, 128 threads.  A (128 x 128 B) and B (N x 128 B) are laid out in shared memory exactly as the product's
// TMA maps write them (K-major, SWIZZLE_128B); the probe computes D = A * B^T three ways and checks each against the
// host: (0) A from shared memory (the product's SS form), (1) A copied smem -> TMEM with tcgen05.cp.128x256b (one
// copy per 32-byte K step, same descriptor the SS form would use for that step), (2) A written to TMEM from registers
// with tcgen05.st.32x32b.x32 (thread = row, register j = bytes 4j..4j+3 of the row).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ts_probe tools/ts_probe.cu && /tmp/ts_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "../scanner_colmap_b200/csrc/ptx.cuh"

using namespace smb::ptx;

constexpr int kN = 224, kACol = 448;

__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p; }" ::"r"(tmem_d),
               "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
               : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};" ::"r"(v[0]),
      "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]),
      "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]),
      "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte (row, col) of a K-major SWIZZLE_128B tile whose base is 1024-byte aligned
__host__ __device__ inline uint32_t sw128(uint32_t row, uint32_t col) {
  return (row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 4) ^ (row & 7)) & 7) << 4) + (col & 15);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* a, const uint8_t* b, int mode, int32_t* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sa = smem_raw + (smem0 - smem_u32(smem_raw));
  uint8_t* sb = sa + 16384;
  for (uint32_t x = threadIdx.x; x < 128 * 128; x += 128) sa[sw128(x >> 7, x & 127)] = a[x];
  for (uint32_t x = threadIdx.x; x < 256 * 128; x += 128) sb[sw128(x >> 7, x & 127)] = (x >> 7) < kN ? b[x] : 0;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc_512(smem_u32(&tmem_base_s));
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tb = tmem_base_s;
  if (mode == 2) {  // A from registers: thread = row
    uint32_t v[32];
    const uint32_t row = threadIdx.x;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = reinterpret_cast<const uint32_t*>(a)[row * 32 + j];
    tmem_st_32x32b_x32(tb + ((warp * 32u) << 16) + kACol, v);
    tmem_wait_st();
    tcgen05_fence_before();
  }
  __syncthreads();
  tcgen05_fence_after();
  if (threadIdx.x == 0) {
    const uint64_t adesc = make_kmajor_sw128_desc(smem0), bdesc = make_kmajor_sw128_desc(smem0 + 16384);
    const uint32_t idesc = make_idesc_u8u8s32(128, kN);
    if (mode == 1)
      for (int k = 0; k < 4; ++k) tmem_cp_128x256b(tb + kACol + k * 8, adesc + k * 2);
    for (int k = 0; k < 4; ++k) {
      if (mode == 0)
        umma_i8(tb, adesc + k * 2, bdesc + k * 2, idesc, k);
      else
        umma_i8_ts(tb, tb + kACol + k * 8, bdesc + k * 2, idesc, k);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < kN; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tb + ((warp * 32u) << 16) + c0, v);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * kN + c0 + j] = (int32_t)v[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc_512(tb);
  }
}

int main() {
  std::vector<uint8_t> a(128 * 128), b(256 * 128);
  uint32_t s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (uint8_t)(s >> 24); };
  for (auto& x : a) x = rnd();
  for (auto& x : b) x = rnd();
  std::vector<int32_t> want(128 * kN);
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < kN; ++j) {
      int32_t d = 0;
      for (int k = 0; k < 128; ++k) d += (int32_t)a[i * 128 + k] * (int32_t)b[j * 128 + k];
      want[i * kN + j] = d;
    }
  uint8_t *da, *db;
  int32_t* dout;
  cudaMalloc(&da, a.size());
  cudaMalloc(&db, b.size());
  cudaMalloc(&dout, want.size() * 4);
  cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice);
  const int smem = 1024 + 16384 + 32768;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[3] = {"SS (A from smem)", "TS, A via tcgen05.cp.128x256b", "TS, A via tcgen05.st.32x32b"};
  int bad_total = 0;
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(dout, 0xFF, want.size() * 4);
    probe_kernel<<<1, 128, smem>>>(da, db, mode, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-34s ERROR %s\n", names[mode], cudaGetErrorString(e));
      return 1;
    }
    std::vector<int32_t> got(want.size());
    cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, first = -1;
    for (size_t x = 0; x < got.size(); ++x)
      if (got[x] != want[x]) {
        if (first < 0) first = (int)x;
        ++bad;
      }
    printf("%-34s mismatches: %d of %zu", names[mode], bad, got.size());
    if (bad) printf("  (first at row %d col %d: got %d want %d)", first / kN, first % kN, got[first], want[first]);
    printf("\n");
    bad_total += mode ? bad : 0;
  }
  return 0;
}
