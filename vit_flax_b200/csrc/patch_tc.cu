// patch_tc.cu -- K1: the patch embedding as ONE kernel: TMA-staged im2col + tcgen05 GEMM + class-token concat +
// positional-embedding add (vit.py:146-153), with the LayerNorm-fold outputs of the first PreNorm riding along.
//
//   x[b*T + cls + t, :] = patch(b, t) W + bias + pos[cls + t]          x[b*T, :] = cls + pos[0]
//   patch(b, t)[(p1*pw + p2)*C + c] = img[b, hh*ph + p1, ww*pw + p2, c],  t = hh*gw + ww      (NHWC images, fp32)
//
// The rearrange of vit.py:146 never exists in memory.  For a fixed p1 the (p2, c) run of a patch is pw*C contiguous
// floats of one image row, so the image is described to TMA as the 5-D tensor [batch, gh, ph, gw, pw*C] (depth = patch
// row, height = row inside the patch, width = patch column, channels = the run) and the patchify is a convolution whose
// filter spans all of H and nothing else: ONE im2col-mode TMA load (cuTensorMapEncodeIm2col;
// cp.async.bulk.tensor.5d...im2col, filter tap h_off = p1) brings the p1-th row of 128 consecutive patches
// -- across patch rows and across images -- as a dense [128 x pw*C] fp32 tile.  K block p1 of the GEMM is that row,
// padded to 64 columns (the packed weight Wt' [D, ph*64] has zero columns there: +33 % MMA work on 0.7 % of the
// forward's FLOPs buys k-blocks that are whole TMA boxes).
//
// One CTA per 128 x 256 output tile, persistent, 16 warps:
//   warp 0  lane 0 : TMA producer: im2col box -> fp32 staging, Wt' box -> B stage            (3-stage ring)
//   warp 1  lane 0 : tcgen05.mma issuer (M 128, N 256, K 16 x 4 per stage), accumulators in TMEM (2 x 256 columns)
//   warp 2         : TMEM allocator
//   warps 4..7     : converters: fp32 staging -> 16-bit, 128-byte-swizzled K-major A stage (tcgen05 has no fp32 MMA and
//                    TMA does not convert, so this hop through registers is the price of reading the fp32 pixels once)
//   warps 8..15    : epilogue, two groups of 4 warps (each owns half of the tile's columns): tcgen05.ld -> + bias + pos ->
//                    fp32 x rows (token placement is a row remap), the class-token rows, and for the LayerNorm fold the
//                    16-bit copy of x and the per-row partial (sum, sum of squares)  (gemm_tc.cu header)
// Eligible when pw*C*4 is a multiple of 16 bytes (TMA stride rule) and pw*C <= 64: every /16 configuration.  ViT-H/14
// (14*3*4 = 168 bytes) and patch 32 keep the patchify kernel + TOKENS GEMM.
#include "common.h"
#include "ptx.cuh"

namespace vb {
namespace {

constexpr int PM = 128, PN = 256, PK = 64;      // tile rows (patches), tile columns, k-block (one p1 row, padded)
constexpr int PSTAGES = 3;
constexpr int P_A_BYTES = PM * PK * 2;          // 16 KB
constexpr int P_B_BYTES = PN * PK * 2;          // 32 KB
constexpr int P_THREADS = 512;
constexpr int P_NUM_CONV = 4, P_NUM_EPI = 8;

__host__ __device__ constexpr int stage_f32_bytes(int run) { return PM * run * 4; }

template <int kDT, bool kLn>
__global__ void __launch_bounds__(P_THREADS, 1)
patch_embed_im2col_kernel(const __grid_constant__ CUtensorMap tmImg,   // im2col map over the images
                          const __grid_constant__ CUtensorMap tmW,     // Wt' [D, ph*64], box 256 x 64
                          const float* __restrict__ bias, const float* __restrict__ pos, const float* __restrict__ cls,
                          float* __restrict__ x, int M /* batch*Np */, int D, int Np, int gw, int ph, int run, int cls_off,
                          LnFold ln) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stage_bytes_f32 = stage_f32_bytes(run);
  const uint32_t sA = base;
  const uint32_t sB = sA + PSTAGES * P_A_BYTES;
  const uint32_t sF = sB + PSTAGES * P_B_BYTES;                        // fp32 staging, PSTAGES x [128, run]
  const uint32_t bars = sF + PSTAGES * uint32_t((stage_bytes_f32 + 1023) & ~1023);
  auto st_full = [&](int s) { return bars + 8u * s; };                 // im2col box landed
  auto st_empty = [&](int s) { return bars + 8u * (PSTAGES + s); };    // converters are done with the staging buffer
  auto ab_full = [&](int s) { return bars + 8u * (2 * PSTAGES + s); }; // B landed + A converted
  auto ab_empty = [&](int s) { return bars + 8u * (3 * PSTAGES + s); };// the stage's MMAs have retired
  auto tfull = [&](int a) { return bars + 8u * (4 * PSTAGES + a); };
  auto tempty = [&](int a) { return bars + 8u * (4 * PSTAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (4 * PSTAGES + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + PM - 1) / PM, n_tiles = (D + PN - 1) / PN;
  const int num_tiles = m_tiles * n_tiles;
  const int T = Np + cls_off;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmImg); prefetch_tmap(&tmW); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PSTAGES; ++s) {
      mbar_init(st_full(s), 1);
      mbar_init(st_empty(s), P_NUM_CONV);
      mbar_init(ab_full(s), 1 + P_NUM_CONV);
      mbar_init(ab_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), P_NUM_EPI); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  if (warp >= 4 && warp < 8) {
    // the pad columns [run, 64) of every A stage stay zero for the whole kernel: clear the stages once
    for (uint32_t i = (threadIdx.x - 128) * 16u; i < PSTAGES * P_A_BYTES; i += 128u * 16u) st_shared_v4(sA + i, 0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int p0 = m_blk * PM;                        // first patch of the tile
        const int b0 = p0 / Np, r0 = p0 - b0 * Np;
        const int hh0 = r0 / gw, ww0 = r0 - hh0 * gw;
        for (int kb = 0; kb < ph; ++kb) {
          mbar_wait(st_empty(stage), phase ^ 1u);
          mbar_wait(ab_empty(stage), phase ^ 1u);
          mbar_arrive_expect_tx(st_full(stage), uint32_t(stage_bytes_f32));
          tma_load_im2col_5d(sF + stage * uint32_t((stage_bytes_f32 + 1023) & ~1023), &tmImg, st_full(stage),
                             0, ww0, 0, hh0, b0, 0, uint16_t(kb), 0);
          mbar_arrive_expect_tx(ab_full(stage), P_B_BYTES);
          tma_load_2d(sB + stage * P_B_BYTES, &tmW, ab_full(stage), kb * PK, n_blk * PN);
          if (++stage == PSTAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_16(PM, PN, kDT == DT_F16 ? 0 : 1);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PN;
        for (int kb = 0; kb < ph; ++kb) {
          mbar_wait(ab_full(stage), phase);
          tc_fence_after();
          const uint32_t a0 = sA + stage * P_A_BYTES, b0 = sB + stage * P_B_BYTES;
#pragma unroll
          for (int k = 0; k < PK / 16; ++k)
            umma_bf16_ss<1>(d_tmem, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), idesc,
                            (kb != 0 || k != 0) ? 1u : 0u);
          umma_commit(ab_empty(stage));
          if (++stage == PSTAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ===================== converters: fp32 staging -> 16-bit swizzled A stage =====================
    const int ct = threadIdx.x - 128;                     // 0..127
    const int run4 = run >> 2;                            // float4 per patch row of the staging tile
    const int total4 = PM * run4;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < ph; ++kb) {
        mbar_wait(st_full(stage), phase);                 // (the producer waited for ab_empty before loading: A[stage] is free)
        const uint32_t f0 = sF + stage * uint32_t((stage_bytes_f32 + 1023) & ~1023), a0 = sA + stage * P_A_BYTES;
        for (int i = ct; i < total4; i += 128) {          // dense float4 reads: conflict-free
          const float4 v = ld_shared_v4f(f0 + uint32_t(i) * 16u);
          const int row = i / run4, c4 = i - row * run4;
          const uint32_t dst = a0 + uint32_t(row) * 128u + (uint32_t((c4 >> 1) ^ (row & 7)) << 4) + uint32_t(c4 & 1) * 8u;
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(pack2<kDT>(v.x, v.y)), "r"(pack2<kDT>(v.z, v.w)) : "memory");
        }
        fence_proxy_async_smem();                         // generic-proxy writes -> the MMA's async-proxy reads
        __syncwarp();
        if (lane == 0) { mbar_arrive(ab_full(stage)); mbar_arrive(st_empty(stage)); }
        if (++stage == PSTAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue =====================
    const int ew = warp - 8, q = ew & 3, grp = ew >> 2;
    const int lrow = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int p = m_blk * PM + lrow;                    // patch index of this thread's row
      const bool row_ok = p < M;
      const int b = p / Np, t = p - b * Np;
      const int64_t orow = int64_t(b) * T + cls_off + t;  // token row (vit.py:151-152: the class token sits in front)
      const bool do_cls = row_ok && cls_off == 1 && t == 0;
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * PN);
      float st1 = 0.f, st2 = 0.f, ct1 = 0.f, ct2 = 0.f;   // partial row sums: this row, and the class-token row it may own
#pragma unroll 1
      for (int chunk = 0; chunk < PN / 64; ++chunk) {     // each group owns half of the tile's columns
        const int n0 = n_blk * PN + grp * (PN / 2) + chunk * 32;
        if (n0 >= D) break;
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + uint32_t(grp * (PN / 2) + chunk * 32), r);
        tmem_ld_wait();
        if (row_ok) {
          float* xrow = x + orow * D;
          const float* prow = pos + int64_t(cls_off + t) * D;
          uint16_t* x16row = kLn ? static_cast<uint16_t*>(ln.x16) + orow * D : nullptr;
          uint32_t h0 = 0, h1 = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int nb = n0 + j * 4;
            if (nb < D) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + nb));
              const float4 p4 = __ldg(reinterpret_cast<const float4*>(prow + nb));
              float4 o;
              o.x = __uint_as_float(r[j * 4 + 0]) + b4.x + p4.x;
              o.y = __uint_as_float(r[j * 4 + 1]) + b4.y + p4.y;
              o.z = __uint_as_float(r[j * 4 + 2]) + b4.z + p4.z;
              o.w = __uint_as_float(r[j * 4 + 3]) + b4.w + p4.w;
              *reinterpret_cast<float4*>(xrow + nb) = o;
              if constexpr (kLn) {
                st1 += (o.x + o.y) + (o.z + o.w);
                st2 = fmaf(o.x, o.x, fmaf(o.y, o.y, fmaf(o.z, o.z, fmaf(o.w, o.w, st2))));
                if ((j & 1) == 0) { h0 = pack2<kDT>(o.x, o.y); h1 = pack2<kDT>(o.z, o.w); }
                else *reinterpret_cast<uint4*>(x16row + nb - 4) = make_uint4(h0, h1, pack2<kDT>(o.x, o.y), pack2<kDT>(o.z, o.w));
              }
            }
          }
          if (do_cls) {
            // the class token in front of image b (vit.py:151-153): row b*T = cls + pos[0], written by the thread that
            // owns the image's first patch row
            float* crow = x + int64_t(b) * T * D;
            uint16_t* c16row = kLn ? static_cast<uint16_t*>(ln.x16) + int64_t(b) * T * D : nullptr;
            uint32_t g0 = 0, g1 = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int nb = n0 + j * 4;
              if (nb < D) {
                const float4 c4 = __ldg(reinterpret_cast<const float4*>(cls + nb));
                const float4 q4 = __ldg(reinterpret_cast<const float4*>(pos + nb));
                const float4 z = make_float4(c4.x + q4.x, c4.y + q4.y, c4.z + q4.z, c4.w + q4.w);
                *reinterpret_cast<float4*>(crow + nb) = z;
                if constexpr (kLn) {
                  ct1 += (z.x + z.y) + (z.z + z.w);
                  ct2 = fmaf(z.x, z.x, fmaf(z.y, z.y, fmaf(z.z, z.z, fmaf(z.w, z.w, ct2))));
                  if ((j & 1) == 0) { g0 = pack2<kDT>(z.x, z.y); g1 = pack2<kDT>(z.z, z.w); }
                  else *reinterpret_cast<uint4*>(c16row + nb - 4) = make_uint4(g0, g1, pack2<kDT>(z.x, z.y), pack2<kDT>(z.z, z.w));
                }
              }
            }
          }
        }
      }
      if constexpr (kLn) {
        // slot (n_blk, group) of the row's statistics; when the consumers expect more slots than this kernel's 256-column
        // tiles fill (small batches: the other producers use 64-column tiles, 4x the slots) the extra ones are zeroed
        if (row_ok) {
          const int own = 2 * n_tiles;
          for (int sl = n_blk * 2 + grp, k = 0; sl < ln.slots; sl += own, ++k) {
            ln.stats[orow * ln.slots + sl] = k == 0 ? make_float2(st1, st2) : make_float2(0.f, 0.f);
            if (do_cls) ln.stats[int64_t(b) * T * ln.slots + sl] = k == 0 ? make_float2(ct1, ct2) : make_float2(0.f, 0.f);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// W fp32 [ph*run, D] (Flax kernel, feature (p1*pw + p2)*C + c) -> Wt' 16-bit [D, ph*64]: column p1*64 + j = feature p1*run + j
template <int kDT>
__global__ void __launch_bounds__(256)
pack_weight_im2col_kernel(const float* __restrict__ W, uint16_t* __restrict__ Wt, int D, int ph, int run) {
  const int64_t total = int64_t(D) * ph * 64;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int n = int(i / (int64_t(ph) * 64)), kk = int(i - int64_t(n) * ph * 64);
    const int p1 = kk >> 6, j = kk & 63;
    Wt[i] = j < run ? cvt16<kDT>(W[(int64_t(p1) * run + j) * D + n]) : uint16_t(0);
  }
}

template <int kDT, bool kLn>
int launch_t(cudaStream_t st, const CUtensorMap& tmImg, const CUtensorMap& tmW, const float* bias, const float* pos,
             const float* cls, float* x, int M, int D, int Np, int gw, int ph, int run, int cls_off, const LnFold& ln) {
  const int smem = PSTAGES * (P_A_BYTES + P_B_BYTES + ((stage_f32_bytes(run) + 1023) & ~1023)) + 1024 + 256;
  static PerDevice<bool> configured_on;
  if (bool& configured = configured_on.here(); !configured) {
    VB_CUDA(cudaFuncSetAttribute(patch_embed_im2col_kernel<kDT, kLn>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  if (smem > 227 * 1024) return fail(VITB200_ERR_UNSUPPORTED, "patch_embed_im2col: stage does not fit shared memory");
  const int tiles = ceil_div(M, PM) * ceil_div(D, PN);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  VB_CUDA(launch_kernel(patch_embed_im2col_kernel<kDT, kLn>, dim3(grid), dim3(P_THREADS), size_t(smem), st, 1, tmImg, tmW,
                        bias, pos, cls, x, M, D, Np, gw, ph, run, cls_off, ln));
  VB_LAUNCH_CHECK("patch_embed_im2col_kernel");
  return 0;
}

}  // namespace

bool patch_im2col_supported(int pw, int channels, int dim) {
  const int run = pw * channels;
  return run <= 64 && (run * 4) % 16 == 0 && dim % 8 == 0;   // (+ ph <= 16, checked where the map is encoded)
}

int launch_pack_weight_im2col(cudaStream_t st, const float* W, void* Wt, int D, int ph, int run, int dtype) {
  const int64_t total = int64_t(D) * ph * 64;
  const int grid = int(std::min<int64_t>((total + 255) / 256, int64_t(sm_count()) * 16));
  if (dtype == DT_BF16) pack_weight_im2col_kernel<DT_BF16><<<grid, 256, 0, st>>>(W, static_cast<uint16_t*>(Wt), D, ph, run);
  else if (dtype == DT_F16) pack_weight_im2col_kernel<DT_F16><<<grid, 256, 0, st>>>(W, static_cast<uint16_t*>(Wt), D, ph, run);
  else return fail(VITB200_ERR_INVALID, "pack_weight_im2col: dtype must be bf16 or fp16");
  VB_LAUNCH_CHECK("pack_weight_im2col_kernel");
  return 0;
}

int launch_patch_embed_im2col(cudaStream_t st, const CUtensorMap& tmImg, const CUtensorMap& tmW, const float* bias,
                              const float* pos, const float* cls, float* x, int batch, int Np, int gw, int ph, int pw,
                              int channels, int D, int cls_off, int dtype, const LnFold* ln) {
  if (batch <= 0 || Np <= 0 || D <= 0) return fail(VITB200_ERR_INVALID, "patch_embed_im2col: empty problem");
  if (!patch_im2col_supported(pw, channels, D))
    return fail(VITB200_ERR_UNSUPPORTED, "patch_embed_im2col: needs pw*C <= 64 floats, a multiple of 16 bytes");
  if (!bias || !pos || !x || (cls_off == 1 && !cls)) return fail(VITB200_ERR_INVALID, "patch_embed_im2col: null pointer");
  const int run = pw * channels;
  const int M = batch * Np;
#define VB_PATCH(DT)                                                                                                        \
  do {                                                                                                                      \
    if (ln) return launch_t<DT, true>(st, tmImg, tmW, bias, pos, cls, x, M, D, Np, gw, ph, run, cls_off, *ln);              \
    return launch_t<DT, false>(st, tmImg, tmW, bias, pos, cls, x, M, D, Np, gw, ph, run, cls_off, LnFold());                \
  } while (0)
  if (dtype == DT_BF16) VB_PATCH(DT_BF16);
  if (dtype == DT_F16) VB_PATCH(DT_F16);
#undef VB_PATCH
  return fail(VITB200_ERR_INVALID, "patch_embed_im2col: dtype must be bf16 or fp16");
}

}  // namespace vb
