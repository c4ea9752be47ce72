// gemm_tc.cu -- K3: persistent, warp-specialised tcgen05 16-bit (bf16 or fp16 operands) GEMM with fused
// epilogues.  acc[M,N] = A[M,K] x Wt[N,K]^T, fp32 accumulation in TMEM.
//
// Replaces every flax `nn.Dense` on the hot path (vit.py:48,51,68,82,147,165)
// together with what follows it in the reference: bias add, tanh-GELU
// (vit.py:49), the Residual add (vit.py:39), cls/pos-embedding placement
// (vit.py:151-153).
//
// Structure (one CTA per SM, 12 warps):
//   warp 0  lane 0 : TMA producer  (A box 128x64, Wt box 256x64, SWIZZLE_128B)
//   warp 1  lane 0 : tcgen05.mma issuer, M=128 N=256 K=16 x4 per smem stage
//   warp 2         : TMEM allocator (512 columns = 2 accumulator stages)
//   warps 4..11    : epilogue: tcgen05.ld 32x32b -> regs -> fused math -> global
// Pipelines: smem ring (4 stages, full/empty mbarriers) between TMA and MMA;
// TMEM double buffer (tfull/tempty mbarriers) between MMA and epilogue, so the
// epilogue of tile i overlaps the MMAs of tile i+1.
#include "common.h"
#include "ptx.cuh"

namespace vb {

namespace {

constexpr int BM = GEMM_BM, BN = GEMM_BN, BK = GEMM_BK;
constexpr int STAGES = 4;
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_BYTES = BN * BK * 2;            // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // 48 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + NUM_EPI_WARPS * 32;   // 384
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ float gelu_tanh_fast(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))  -- nn.gelu default (vit.py:49)
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float u = k0 * x * fmaf(k1 * x, x, 1.0f);
  return 0.5f * x * (1.0f + tanh_approx(u));
}

template <int kEpi, int kDT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                    const __grid_constant__ CUtensorMap tmB,
                    const float* __restrict__ bias, void* __restrict__ Cout,
                    int M, int N, int K, const float* __restrict__ aux, int tpi) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES * A_BYTES;
  const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
  // barrier layout: full[4] empty[4] tfull[2] tempty[2] | tmem ptr
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  // generic pointer to the tmem slot for reading it back
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BM - 1) / BM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), NUM_EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_2d(sA + stage * A_BYTES, &tmA, full_bar(stage), kb * BK, m_blk * BM);
          tma_load_2d(sB + stage * B_BYTES, &tmB, full_bar(stage), kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_16(BM, BN, kDT == DT_F16 ? 0 : 1);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a0 = sA + stage * A_BYTES, b0 = sB + stage * B_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_bf16_ss<1>(d_tmem, umma_desc_k_sw128(a0 + k * UMMA_K * 2),
                            umma_desc_k_sw128(b0 + k * UMMA_K * 2), idesc,
                            (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));          // smem slot free once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));              // accumulator ready for the epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;
    const int q = ew & 3;        // TMEM lane quarter this warp may access (== warp % 4)
    const int hf = ew >> 2;      // which 128-column half of the 256-column accumulator
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < M;
      int64_t out_row = row;
      const float* pos_row = nullptr;
      if constexpr (kEpi == VITB200_EPI_PATCH_F32) {
        const int b = row / tpi, t = row - b * tpi;
        out_row = int64_t(b) * (tpi + 1) + 1 + t;
        pos_row = aux + int64_t(1 + t) * N;
      }
#pragma unroll 1
      for (int chunk = 0; chunk < 4; ++chunk) {
        const int col0 = hf * 128 + chunk * 32;
        const int n0 = n_blk * BN + col0;
        if (n0 >= N) break;                         // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * BN + col0), r);
        tmem_ld_wait();
        if (row_ok) {
          if constexpr (kEpi == VITB200_EPI_STORE_16 || kEpi == VITB200_EPI_BIAS_GELU_16) {
            uint16_t* crow = reinterpret_cast<uint16_t*>(Cout) + out_row * N + n0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (n0 + j * 8 < N) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[j * 8 + e]);
                if constexpr (kEpi == VITB200_EPI_BIAS_GELU_16) {
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + j * 8));
                  const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + j * 8 + 4));
                  v[0] = gelu_tanh_fast(v[0] + b0.x); v[1] = gelu_tanh_fast(v[1] + b0.y);
                  v[2] = gelu_tanh_fast(v[2] + b0.z); v[3] = gelu_tanh_fast(v[3] + b0.w);
                  v[4] = gelu_tanh_fast(v[4] + b1.x); v[5] = gelu_tanh_fast(v[5] + b1.y);
                  v[6] = gelu_tanh_fast(v[6] + b1.z); v[7] = gelu_tanh_fast(v[7] + b1.w);
                }
                uint4 o;
                o.x = pack2<kDT>(v[0], v[1]); o.y = pack2<kDT>(v[2], v[3]);
                o.z = pack2<kDT>(v[4], v[5]); o.w = pack2<kDT>(v[6], v[7]);
                *reinterpret_cast<uint4*>(crow + j * 8) = o;
              }
            }
          } else {
            float* crow = reinterpret_cast<float*>(Cout) + out_row * N + n0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (n0 + j * 4 < N) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + j * 4));
                float4 o;
                o.x = __uint_as_float(r[j * 4 + 0]) + b4.x;
                o.y = __uint_as_float(r[j * 4 + 1]) + b4.y;
                o.z = __uint_as_float(r[j * 4 + 2]) + b4.z;
                o.w = __uint_as_float(r[j * 4 + 3]) + b4.w;
                if constexpr (kEpi == VITB200_EPI_BIAS_RESID_F32) {
                  const float4 x4 = *reinterpret_cast<const float4*>(crow + j * 4);
                  o.x += x4.x; o.y += x4.y; o.z += x4.z; o.w += x4.w;
                }
                if constexpr (kEpi == VITB200_EPI_PATCH_F32) {
                  const float4 p4 = __ldg(reinterpret_cast<const float4*>(pos_row + n0 + j * 4));
                  o.x += p4.x; o.y += p4.y; o.z += p4.z; o.w += p4.w;
                }
                *reinterpret_cast<float4*>(crow + j * 4) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, TMEM_COLS);
  }
}

template <int kEpi, int kDT>
int launch_one(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB,
               const float* bias, void* C, int M, int N, int K, const float* aux, int tpi) {
  static bool configured = false;   // per-process; attribute is per-function, device-agnostic
  if (!configured) {
    VB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<kEpi, kDT>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int tiles = ceil_div(M, BM) * ceil_div(N, BN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  gemm_tc_kernel<kEpi, kDT><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, bias, C, M, N, K,
                                                                       aux, tpi);
  VB_LAUNCH_CHECK("gemm_tc_kernel");
  return 0;
}

template <int kDT>
int dispatch_epi(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB,
                 const float* bias, void* C, int M, int N, int K, int epilogue, const float* aux,
                 int tpi) {
  switch (epilogue) {
    case VITB200_EPI_STORE_16:
      return launch_one<VITB200_EPI_STORE_16, kDT>(stream, tmA, tmB, bias, C, M, N, K, aux, tpi);
    case VITB200_EPI_BIAS_GELU_16:
      return launch_one<VITB200_EPI_BIAS_GELU_16, kDT>(stream, tmA, tmB, bias, C, M, N, K, aux, tpi);
    case VITB200_EPI_BIAS_RESID_F32:
      return launch_one<VITB200_EPI_BIAS_RESID_F32, kDT>(stream, tmA, tmB, bias, C, M, N, K, aux, tpi);
    case VITB200_EPI_BIAS_F32:
      return launch_one<VITB200_EPI_BIAS_F32, kDT>(stream, tmA, tmB, bias, C, M, N, K, aux, tpi);
    case VITB200_EPI_PATCH_F32:
      if (aux == nullptr || tpi <= 0)
        return fail(VITB200_ERR_INVALID, "gemm_tc: PATCH epilogue needs pos_embedding and tokens");
      return launch_one<VITB200_EPI_PATCH_F32, kDT>(stream, tmA, tmB, bias, C, M, N, K, aux, tpi);
    default:
      return fail(VITB200_ERR_INVALID, "gemm_tc: unknown epilogue");
  }
}

}  // namespace

int launch_gemm_tc(cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmB,
                   const float* bias, void* C, int M, int N, int K, int epilogue,
                   const float* aux, int tpi, int dtype) {
  if (M <= 0 || N <= 0 || K <= 0) return fail(VITB200_ERR_INVALID, "gemm_tc: empty problem");
  if ((N % 8) != 0 || (K % 8) != 0)
    return fail(VITB200_ERR_INVALID, "gemm_tc: N and K must be multiples of 8");
  if (epilogue != VITB200_EPI_STORE_16 && bias == nullptr)
    return fail(VITB200_ERR_INVALID, "gemm_tc: epilogue needs a bias");
  if (dtype == DT_BF16)
    return dispatch_epi<DT_BF16>(stream, tmA, tmB, bias, C, M, N, K, epilogue, aux, tpi);
  if (dtype == DT_F16)
    return dispatch_epi<DT_F16>(stream, tmA, tmB, bias, C, M, N, K, epilogue, aux, tpi);
  return fail(VITB200_ERR_INVALID, "gemm_tc: dtype must be bf16 or fp16");
}

}  // namespace vb
