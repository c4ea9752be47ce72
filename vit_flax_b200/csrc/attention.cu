// attention.cu -- K4 (first generation): fused flash-style attention,
//   out = softmax(Q K^T * 64^-0.5) V   per (image, head)          vit.py:69-79
// reading q/k/v straight out of the to_qkv output [B*T, 3*I] (the jnp.split and
// 'b n (h d) -> b h n d' rearranges of vit.py:69-71 are pure index math here)
// and writing the merged-head layout [B*T, I] of vit.py:79 directly.  The
// [B,h,T,T] score tensor of vit.py:73-75 never exists: S lives in registers,
// softmax is the online (running max / running sum) form.
//
// This generation uses warp-level mma.sync (HMMA) with ldmatrix operands and a
// double-buffered cp.async K/V stream; each warp owns 16 query rows.  The tcgen05/TMEM kernels
// (attention_tc5.cu for T <= 208, attention_tc5m.cu beyond) have replaced it on the hot path;
// it stays reachable with VITB200_ATTENTION=hmma as an independent implementation for A/B tests
// (tests/test_gpu_kernels.py runs both at every shape).
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace vb {

namespace {

constexpr int DH = 64;            // dim_head, vit.py:123
constexpr int KV_BLOCK = 64;      // keys per pipeline stage
constexpr int MAX_WARPS = 16;
constexpr int ROW_BYTES = DH * 2; // 128 B per q/k/v row in shared memory

// 16-byte chunk c of row r lives at chunk (c ^ (r & 7)): conflict-free ldmatrix.
__device__ __forceinline__ uint32_t swz(uint32_t base, int row, int chunk) {
  return base + uint32_t(row) * ROW_BYTES + (uint32_t(chunk ^ (row & 7)) << 4);
}

// rows [row0, row0+nrows) of a [*, ld] bf16 matrix (64 columns starting at col0) -> smem
__device__ __forceinline__ void load_rows_async(uint32_t sbase, const uint16_t* g, int64_t ld,
                                                int row0, int nrows, int row_limit, int tid,
                                                int nthreads) {
  for (int i = tid; i < nrows * 8; i += nthreads) {
    const int r = i >> 3, c = i & 7;
    const int grow = row0 + r;
    const bool ok = grow < row_limit;
    const uint16_t* src = g + int64_t(ok ? grow : 0) * ld + c * 8;
    cp_async16(swz(sbase, r, c), src, ok);
  }
}

template <int kDT>
__global__ void __launch_bounds__(MAX_WARPS * 32, 1)
attention_tc_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out,
                      int T, int heads) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int nwarps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tg = lane & 3;
  const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads;
  const int inner = heads * DH;
  const int64_t ld = 3 * int64_t(inner);
  const uint16_t* qbase = qkv + int64_t(b) * T * ld + h * DH;
  const uint16_t* kbase = qbase + inner;
  const uint16_t* vbase = qbase + 2 * inner;

  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + uint32_t(nwarps) * 16 * ROW_BYTES;
  const uint32_t sV = sK + 2 * KV_BLOCK * ROW_BYTES;

  const int q_cta0 = blockIdx.x * nwarps * 16;
  const int q0 = q_cta0 + warp * 16;
  const int nkb = (T + KV_BLOCK - 1) / KV_BLOCK;

  // prologue: Q tile + KV block 0
  load_rows_async(sQ, qbase, ld, q_cta0, nwarps * 16, T, threadIdx.x, blockDim.x);
  load_rows_async(sK, kbase, ld, 0, KV_BLOCK, T, threadIdx.x, blockDim.x);
  load_rows_async(sV, vbase, ld, 0, KV_BLOCK, T, threadIdx.x, blockDim.x);
  cp_async_commit();

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const float sl2 = 0.125f * 1.4426950408889634f;   // dim_head^-0.5 * log2(e)
  const bool warp_active = q0 < T;

  for (int kb = 0; kb < nkb; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nkb) {
      load_rows_async(sK + (buf ^ 1) * KV_BLOCK * ROW_BYTES, kbase, ld, (kb + 1) * KV_BLOCK,
                      KV_BLOCK, T, threadIdx.x, blockDim.x);
      load_rows_async(sV + (buf ^ 1) * KV_BLOCK * ROW_BYTES, vbase, ld, (kb + 1) * KV_BLOCK,
                      KV_BLOCK, T, threadIdx.x, blockDim.x);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (warp_active) {
      if (kb == 0) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const int mi = lane >> 3, r = lane & 7;
          const int row = warp * 16 + (mi & 1) * 8 + r;
          ldmatrix_x4(swz(sQ, row, ks * 2 + (mi >> 1)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
        }
      }
      const uint32_t kS = sK + buf * KV_BLOCK * ROW_BYTES;
      const uint32_t vS = sV + buf * KV_BLOCK * ROW_BYTES;
      const int keys_here = min(KV_BLOCK, T - kb * KV_BLOCK);
      const int npairs = (keys_here + 15) >> 4;        // 16-key groups with any valid key

      // ---- S = Q K^T ----
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (p < npairs) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const int mi = lane >> 3, r = lane & 7;
            const int key = p * 16 + (mi >> 1) * 8 + r;
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(swz(kS, key, ks * 2 + (mi & 1)), b0, b1, b2, b3);
            mma_16816<kDT>(s[2 * p], qf[ks], b0, b1);
            mma_16816<kDT>(s[2 * p + 1], qf[ks], b2, b3);
          }
        }
      }
      // ---- mask + online softmax ----
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c0 = nt * 8 + 2 * tg;
        if (c0 >= keys_here) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (c0 + 1 >= keys_here) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
        mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
      }
      float corr[2], mneg[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float m_new = fmaxf(m_run[r], mx[r]);
        corr[r] = ex2_approx((m_run[r] - m_new) * sl2);
        m_run[r] = m_new;
        mneg[r] = -m_new * sl2;
        l_run[r] *= corr[r];
      }
      uint32_t pf[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = ex2_approx(fmaf(s[nt][0], sl2, mneg[0]));
        const float p1 = ex2_approx(fmaf(s[nt][1], sl2, mneg[0]));
        const float p2 = ex2_approx(fmaf(s[nt][2], sl2, mneg[1]));
        const float p3 = ex2_approx(fmaf(s[nt][3], sl2, mneg[1]));
        l_run[0] += p0 + p1;
        l_run[1] += p2 + p3;
        const int j = nt >> 1;
        if ((nt & 1) == 0) { pf[j][0] = pack2<kDT>(p0, p1); pf[j][1] = pack2<kDT>(p2, p3); }
        else               { pf[j][2] = pack2<kDT>(p0, p1); pf[j][3] = pack2<kDT>(p2, p3); }
      }
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        o[dt][0] *= corr[0]; o[dt][1] *= corr[0];
        o[dt][2] *= corr[1]; o[dt][3] *= corr[1];
      }
      // ---- O += P V ----
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < npairs) {
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {
            const int mi = lane >> 3, r = lane & 7;
            const int key = j * 16 + (mi & 1) * 8 + r;
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(swz(vS, key, dp * 2 + (mi >> 1)), b0, b1, b2, b3);
            mma_16816<kDT>(o[2 * dp], pf[j], b0, b1);
            mma_16816<kDT>(o[2 * dp + 1], pf[j], b2, b3);
          }
        }
      }
    }
    __syncthreads();   // all warps done with buf before it is refilled
  }

  if (warp_active) {
    float inv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float l = l_run[r];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      inv[r] = 1.f / l;
    }
    // stage the 16x64 tile in this warp's own Q rows, then 16-byte coalesced stores
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      const int row_a = warp * 16 + g, row_b = row_a + 8;
      const uint32_t off = uint32_t(tg) * 4;      // 2 bf16 = 4 bytes inside the 16 B chunk
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(swz(sQ, row_a, dt) + off),
                   "r"(pack2<kDT>(o[dt][0] * inv[0], o[dt][1] * inv[0])) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(swz(sQ, row_b, dt) + off),
                   "r"(pack2<kDT>(o[dt][2] * inv[1], o[dt][3] * inv[1])) : "memory");
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane;
      const int r = idx >> 3, c = idx & 7;
      const int qrow = q0 + r;
      if (qrow < T) {
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(swz(sQ, warp * 16 + r, c)) : "memory");
        *reinterpret_cast<uint4*>(out + (int64_t(b) * T + qrow) * inner + h * DH + c * 8) = v;
      }
    }
  }
}

}  // namespace

template <int kDT>
static int launch_attention_t(cudaStream_t stream, const uint16_t* qkv, uint16_t* out, int batch,
                              int T, int heads) {
  const int nq16 = ceil_div(T, 16);
  const int ctas = ceil_div(nq16, MAX_WARPS);
  const int nwarps = ceil_div(nq16, ctas);
  const size_t smem = size_t(nwarps) * 16 * ROW_BYTES + 4 * size_t(KV_BLOCK) * ROW_BYTES;
  static PerDevice<bool> configured_on;   // the smem opt-in is per (function, device)
  if (bool& configured = configured_on.here(); !configured) {
    VB_CUDA(cudaFuncSetAttribute(attention_tc_kernel<kDT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 MAX_WARPS * 16 * ROW_BYTES + 4 * KV_BLOCK * ROW_BYTES));
    configured = true;
  }
  dim3 grid(ctas, batch * heads);
  attention_tc_kernel<kDT><<<grid, nwarps * 32, smem, stream>>>(qkv, out, T, heads);
  VB_LAUNCH_CHECK("attention_tc_kernel");
  return 0;
}

int launch_attention_tc(cudaStream_t stream, const void* qkv, void* out, int batch, int T, int heads,
                        int dtype, float* lse) {
  if (batch <= 0 || T <= 0 || heads <= 0)
    return fail(VITB200_ERR_INVALID, "attention_tc: empty problem");
  // VITB200_ATTENTION=hmma forces the mma.sync generation (A/B tests); default: the tcgen05
  // kernels -- one key block when T <= 208, streamed key blocks with an online softmax beyond.
  const char* force = getenv("VITB200_ATTENTION");
  const bool want_hmma = force && force[0] == 'h';
  if (!want_hmma && attention_tc5_supports(T))
    return launch_attention_tc5(stream, qkv, out, batch, T, heads, dtype, lse);
  if (!want_hmma) return launch_attention_tc5m(stream, qkv, out, batch, T, heads, dtype, lse);
  if (lse != nullptr) return fail(VITB200_ERR_UNSUPPORTED, "attention_tc: the row log-sum-exp output needs a tcgen05 kernel");
  if (int64_t(batch) * heads > 65535)
    return fail(VITB200_ERR_INVALID, "attention_tc: batch*heads exceeds grid.y limit; chunk the batch");
  if (dtype == DT_BF16)
    return launch_attention_t<DT_BF16>(stream, static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out), batch, T, heads);
  if (dtype == DT_F16)
    return launch_attention_t<DT_F16>(stream, static_cast<const uint16_t*>(qkv), static_cast<uint16_t*>(out), batch, T, heads);
  return fail(VITB200_ERR_INVALID, "attention_tc: dtype must be bf16 or fp16");
}

}  // namespace vb
