// Latency / issue rate of SMALL tcgen05.mma chains on B200 (sm_100a) -- the shapes the attention kernels are made of.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../vit_flax_b200/csrc -o mma_chain mma_chain.cu && ./mma_chain
// One CTA per SM (148), one issuing thread; M = 128, cta_group::1, bf16, K = 16 per instruction.  For every case:
//   L instructions issued back to back, then ONE tcgen05.commit; reports cycles from the first issue to (a) the last
//   issue returning ("issue": back-pressure on the issuing thread) and (b) the commit's mbarrier completing ("done").
// Cases: accumulate-chains into one D (what P V and Q K^T are), round-robin over 2 / 4 independent D, A from shared
// memory (SS) or from TMEM (TS), and the chains while 4 / 8 other warps run tcgen05.ld / st / MUFU loops like the softmax passes.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace vb;

struct Case {
  int N, L, accs, ts, ld_warps;
};

template <int N, int L, int ACCS, int TS>
__global__ void __launch_bounds__(512, 1) k(Case c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bars = base + 16384 + 32768;
  const uint32_t done = bars, tmem_slot = bars + 8, stop = bars + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (bars + 8 - smem_u32(smem_raw)));
  volatile uint32_t* stop_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (bars + 16 - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(done, 1); *stop_ptr = 0; fence_barrier_init(); }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc_k = umma_idesc_16(128, N, 1, 0);    // B K-major (Q K^T)
    constexpr uint32_t idesc_mn = umma_idesc_16(128, N, 1, 1);   // B MN-major (P V)
    long long t_issue = 0, t_done = 0;
    for (int rep = 0; rep < 8; ++rep) {                          // the last repetition is reported
      tc_fence_after();
      const long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < L; ++i) {                              // fully unrolled: every operand is a compile-time offset
        const int a = i % ACCS;
        const uint32_t d = tmem_base + 256 + a * 64;             // accumulators at columns 256.. (N <= 64 when ACCS > 1)
        const uint32_t acc = i >= ACCS ? 1u : 0u;
        // TS: 0 = both operands K-major from smem (Q K^T), 1 = A from TMEM, B MN-major (forward P V), 2 = A MN-major two
        // atoms wide + B MN-major (adjoint dV = P^T dO, dK = dS^T Q), 3 = A K-major + B MN-major (adjoint dQ = dS K)
        if (TS == 1) umma_bf16_ts(d, tmem_base + (i % 13) * 8, umma_desc_mn_sw128(sB + (i % 13) * 2048), idesc_mn, acc);
        else if (TS == 2) umma_bf16_ss<1>(d, umma_desc_mn_sw128_wide(sB + (i & 7) * 2048, 16384), umma_desc_mn_sw128(sA + (i & 7) * 2048),
                                          umma_idesc_16(128, N, 1, 1, 1), acc);
        else if (TS == 3) umma_bf16_ss<1>(d, umma_desc_k_sw128(sB + (i >> 2 & 1) * 16384 + (i & 3) * 32), umma_desc_mn_sw128(sA + (i & 7) * 2048),
                                          idesc_mn, acc);
        else umma_bf16_ss<1>(d, umma_desc_k_sw128(sA + (i & 3) * 32), umma_desc_k_sw128(sB + (i & 3) * 32), idesc_k, acc);
      }
      umma_commit(done);
      const long long t1 = clock64();
      mbar_wait(done, rep & 1);
      const long long t2 = clock64();
      t_issue = t1 - t0;
      t_done = t2 - t0;
    }
    *stop_ptr = 1;
    if (blockIdx.x == 0) { out[0] = t_issue; out[1] = t_done; }
  } else if (warp >= 4 && warp < 4 + c.ld_warps) {
    // background warps, like the softmax passes.  c.ts >> 4 selects what they do: 0 = tcgen05.ld over columns [0, 192)
    // of their lane quarter; 1 = ld + tcgen05.st (columns 192..207); 2 = MUFU only (no TMEM); 3 = ld + MUFU + st
    const int mode = c.ts >> 4;   // (the low four bits are the operand form)
    uint32_t r[32];
    float acc = 0.f;
    const uint32_t t_lane = tmem_base + (uint32_t((warp & 3) * 32) << 16);
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = 0x3f000000u + j;
    while (*stop_ptr == 0) {
      for (int cc = 0; cc < 6; ++cc) {
        if (mode != 2) {
          tmem_ld_32x32b_x32p(t_lane + cc * 32, r);
          tmem_ld_wait();
        }
        if (mode >= 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(ex2_approx(__uint_as_float(r[j])));
        }
        if (mode == 1 || mode == 3) tmem_st_32x32b_x16(t_lane + 192, r);
        acc += __uint_as_float(r[lane & 31]);
      }
      if (mode == 1 || mode == 3) tmem_st_wait();
    }
    if (acc == 1234.5f) out[2] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc<1>(tmem_base, 512); }
}

template <int N, int L, int ACCS, int TS>
void run(long long* out, int ld_warps) {
  const int smem = 16384 + 32768 + 1024 + 1024;
  cudaFuncSetAttribute(k<N, L, ACCS, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Case c{N, L, ACCS, (TS & 15) | ((ld_warps >> 8) << 4), ld_warps & 255};
  ld_warps = c.ld_warps;
  k<N, L, ACCS, TS><<<148, 512, smem>>>(c, out);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
  long long h[2];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%5d %4d %5d %3d %9d | %8lld %8lld %10.1f\n", N, L, ACCS, int(TS), ld_warps, h[0], h[1], double(h[1]) / L);
}

int main() {
  long long* out;
  cudaMalloc(&out, 64);
  printf("%5s %4s %5s %3s %9s | %8s %8s %10s\n", "N", "L", "accs", "TS", "ld_warps", "issue", "done", "done/L");
  run<16, 16, 1, 0>(out, 0);
  run<64, 16, 1, 0>(out, 0);
  run<128, 16, 1, 0>(out, 0);
  run<208, 16, 1, 0>(out, 0);
  run<256, 16, 1, 0>(out, 0);
  run<208, 4, 1, 0>(out, 0);
  run<208, 1, 1, 0>(out, 0);
  run<64, 1, 1, 1>(out, 0);
  run<64, 4, 1, 1>(out, 0);
  run<64, 13, 1, 1>(out, 0);
  run<64, 16, 1, 1>(out, 0);
  run<32, 16, 1, 1>(out, 0);
  run<64, 16, 2, 1>(out, 0);
  run<64, 16, 4, 1>(out, 0);
  run<64, 16, 2, 0>(out, 0);
  run<64, 32, 1, 1>(out, 0);
  printf("adjoint shapes: A MN-major wide + B MN-major (TS=2), A K-major + B MN-major (TS=3), 8 and 24 in a row\n");
  run<64, 8, 1, 0>(out, 0);
  run<64, 8, 1, 2>(out, 0);
  run<64, 8, 1, 3>(out, 0);
  run<64, 24, 1, 2>(out, 0);
  run<64, 24, 3, 2>(out, 0);
  run<64, 8, 1, 2>(out, 8 | (3 << 8));
  printf("background warps (ld_warps): tcgen05.ld\n");
  run<64, 13, 1, 1>(out, 4);
  run<64, 13, 1, 1>(out, 8);
  run<208, 4, 1, 0>(out, 8);
  printf("background: ld + st\n");
  run<64, 13, 1, 1>(out, 8 | (1 << 8));
  run<208, 4, 1, 0>(out, 8 | (1 << 8));
  printf("background: MUFU only\n");
  run<64, 13, 1, 1>(out, 8 | (2 << 8));
  run<208, 4, 1, 0>(out, 8 | (2 << 8));
  printf("background: ld + MUFU + st\n");
  run<64, 13, 1, 1>(out, 4 | (3 << 8));
  run<64, 13, 1, 1>(out, 8 | (3 << 8));
  run<208, 4, 1, 0>(out, 8 | (3 << 8));
  return 0;
}
