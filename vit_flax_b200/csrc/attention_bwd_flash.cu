// attention_bwd_flash.cu -- attention adjoint for ANY token count (adjoint of vit.py:69-79): the
// streamed, flash-style form of backward.cu's shared-memory-resident kernel (which needs the whole
// T x T probability matrix of a head on chip and stops at T = 208).
//
//   stats kernel   per (image, head, 64 query rows): lse2_i = log2 sum_j exp2(s_ij), s = q k^T / 8 * log2(e)
//                  (online over 64-key blocks) and D_i = sum_d dO_id O_id
//   main kernel    per (image, head, 64 KEYS): K_j, V_j stay in shared memory, each warp owns 16 keys and keeps
//                  dK, dV for them in registers; for every 64-query tile:
//                      S^T = K_j Q_i^T  ->  P^T = exp2(S^T - lse2_i)            (transposed on purpose: P^T and
//                      dV_j += P^T dO_i                                          dS^T come out of the MMA in the
//                      dP^T = V_j dO_i^T ;  dS^T = P^T o (dP^T - D_i) / 8        layout the next MMA wants as A)
//                      dK_j += dS^T Q_i
//                      dQ_i += dS K_j   via a 64 x 64 staging tile (read back transposed) and fp32 atomics,
//                                       because the key blocks of a head run in different CTAs
//   convert kernel dQ fp32 -> the q third of dqkv (16-bit)
// All matmuls are mma.sync m16n8k16 with ldmatrix operands (the tcgen05 form is future work).
#include <algorithm>

#include "common.h"
#include "ptx.cuh"

namespace vb {

namespace {

constexpr int DH = 64;
constexpr int ROW_BYTES = DH * 2;
constexpr int TILE = 64;                      // query rows per tile == keys per CTA
constexpr int TILE_BYTES = TILE * ROW_BYTES;  // 8 KB

__device__ __forceinline__ uint32_t swz(uint32_t base, int row, int chunk) {
  return base + uint32_t(row) * ROW_BYTES + (uint32_t(chunk ^ (row & 7)) << 4);
}
// rows [row0, row0 + 64) of a [*, ld] 16-bit matrix (64 columns) -> swizzled smem tile; rows >= limit are zero
__device__ __forceinline__ void load_tile(uint32_t sbase, const uint16_t* g, int64_t ld, int row0, int limit) {
  for (int i = threadIdx.x; i < TILE * 8; i += blockDim.x) {
    const int r = i >> 3, c = i & 7;
    const bool ok = row0 + r < limit;
    cp_async16(swz(sbase, r, c), g + int64_t(ok ? row0 + r : 0) * ld + c * 8, ok);
  }
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
template <int kDT>
__device__ __forceinline__ void unpack2(uint32_t v, float& lo, float& hi) {
  lo = to_f32<kDT>(uint16_t(v & 0xFFFFu));
  hi = to_f32<kDT>(uint16_t(v >> 16));
}

// acc[8][4] (16 rows x 64 columns) = A-frags af (16 x 64, the K dimension) . X[64 rows, 64]^T, X row-major in smem
template <int kDT>
__device__ __forceinline__ void mma_abt(float (&acc)[8][4], const uint32_t (&af)[4][4], uint32_t sX, int lane) {
  const int mi = lane >> 3, r8 = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4(swz(sX, p * 16 + (mi >> 1) * 8 + r8, ks * 2 + (mi & 1)), b0, b1, b2, b3);
      mma_16816<kDT>(acc[2 * p], af[ks], b0, b1);
      mma_16816<kDT>(acc[2 * p + 1], af[ks], b2, b3);
    }
  }
}
// acc[8][4] (16 x 64) += A-frags pf (16 x 64 over the rows of X) . X[64 rows, 64], X row-major in smem
template <int kDT>
__device__ __forceinline__ void mma_ab(float (&acc)[8][4], const uint32_t (&pf)[4][4], uint32_t sX, int lane) {
  const int mi = lane >> 3, r8 = lane & 7;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4_trans(swz(sX, j * 16 + (mi & 1) * 8 + r8, dp * 2 + (mi >> 1)), b0, b1, b2, b3);
      mma_16816<kDT>(acc[2 * dp], pf[j], b0, b1);
      mma_16816<kDT>(acc[2 * dp + 1], pf[j], b2, b3);
    }
  }
}

// ------------------------------------------------------------------ statistics
template <int kDT>
__global__ void __launch_bounds__(128)
attn_bwd_stats_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ o_fwd, const uint16_t* __restrict__ d_out,
                      float* __restrict__ lse2, float* __restrict__ dsum, int T, int heads) {
  __shared__ __align__(128) uint8_t smem[2 * TILE_BYTES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3, mi = lane >> 3, r8 = lane & 7;
  const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads, q0 = blockIdx.x * TILE;
  const int inner = heads * DH;
  const int64_t ld = 3 * int64_t(inner);
  const uint16_t* qbase = qkv + int64_t(b) * T * ld + h * DH;
  const uint32_t sQ = smem_u32(smem), sK = sQ + TILE_BYTES;
  load_tile(sQ, qbase, ld, q0, T);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldmatrix_x4(swz(sQ, warp * 16 + (mi & 1) * 8 + r8, ks * 2 + (mi >> 1)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  const float sl2 = 0.125f * 1.4426950408889634f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  for (int k0 = 0; k0 < T; k0 += TILE) {
    __syncthreads();                       // every warp is done with the previous key block
    load_tile(sK, qbase + inner, ld, k0, T);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    float s[8][4];
    mma_abt<kDT>(s, qf, sK, lane);
    float c0m = -INFINITY, c1m = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c0 = k0 + nt * 8 + 2 * tg;
      if (c0 >= T) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
      if (c0 + 1 >= T) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
      c0m = fmaxf(c0m, fmaxf(s[nt][0], s[nt][1]));
      c1m = fmaxf(c1m, fmaxf(s[nt][2], s[nt][3]));
    }
    c0m = fmaxf(c0m, __shfl_xor_sync(0xffffffffu, c0m, 1));
    c0m = fmaxf(c0m, __shfl_xor_sync(0xffffffffu, c0m, 2));
    c1m = fmaxf(c1m, __shfl_xor_sync(0xffffffffu, c1m, 1));
    c1m = fmaxf(c1m, __shfl_xor_sync(0xffffffffu, c1m, 2));
    const float n0 = fmaxf(m0, c0m), n1 = fmaxf(m1, c1m);      // finite: every block holds a key < T
    l0 *= ex2_approx((m0 - n0) * sl2);
    l1 *= ex2_approx((m1 - n1) * sl2);
    m0 = n0;
    m1 = n1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      l0 += ex2_approx((s[nt][0] - m0) * sl2) + ex2_approx((s[nt][1] - m0) * sl2);
      l1 += ex2_approx((s[nt][2] - m1) * sl2) + ex2_approx((s[nt][3] - m1) * sl2);
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // D_i = sum_d dO[i, d] O[i, d]: the quad of a row pair splits the 64 columns
  const int ra = q0 + warp * 16 + g, rb = ra + 8;
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    float x0, x1, y0, y1;
    if (ra < T) {
      const int64_t e = (int64_t(b) * T + ra) * inner + h * DH + nt * 8 + 2 * tg;
      unpack2<kDT>(*reinterpret_cast<const uint32_t*>(d_out + e), x0, x1);
      unpack2<kDT>(*reinterpret_cast<const uint32_t*>(o_fwd + e), y0, y1);
      d0 += x0 * y0 + x1 * y1;
    }
    if (rb < T) {
      const int64_t e = (int64_t(b) * T + rb) * inner + h * DH + nt * 8 + 2 * tg;
      unpack2<kDT>(*reinterpret_cast<const uint32_t*>(d_out + e), x0, x1);
      unpack2<kDT>(*reinterpret_cast<const uint32_t*>(o_fwd + e), y0, y1);
      d1 += x0 * y0 + x1 * y1;
    }
  }
  d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
  d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
  d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
  d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
  if (tg == 0) {
    if (ra < T) { lse2[int64_t(bh) * T + ra] = m0 * sl2 + log2f(l0); dsum[int64_t(bh) * T + ra] = d0; }
    if (rb < T) { lse2[int64_t(bh) * T + rb] = m1 * sl2 + log2f(l1); dsum[int64_t(bh) * T + rb] = d1; }
  }
}

// ------------------------------------------------------------------ main
template <int kDT>
__global__ void __launch_bounds__(128, 3)
attn_bwd_flash_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ d_out, const float* __restrict__ lse2,
                      const float* __restrict__ dsum, uint16_t* __restrict__ dqkv, float* __restrict__ dq_acc, int T, int heads) {
  extern __shared__ __align__(128) uint8_t smem[];     // K, V, dS^T staging, 2 x (Q, dO) tiles, 2 x (lse2, D)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3, mi = lane >> 3, r8 = lane & 7;
  const int bh = blockIdx.y, b = bh / heads, h = bh - b * heads, k0 = blockIdx.x * TILE;
  const int inner = heads * DH;
  const int64_t ld = 3 * int64_t(inner);
  const uint16_t* qbase = qkv + int64_t(b) * T * ld + h * DH;
  const uint16_t* dobase = d_out + int64_t(b) * T * inner + h * DH;
  uint16_t* dqbase = dqkv + int64_t(b) * T * ld + h * DH;
  const uint32_t sK = smem_u32(smem), sV = sK + TILE_BYTES, sT = sV + TILE_BYTES, sQD = sT + TILE_BYTES;
  float* s_stat = reinterpret_cast<float*>(smem + 7 * TILE_BYTES);     // [2][2][TILE]
  auto fetch_q_tile = [&](int q0, int buf) {
    load_tile(sQD + buf * 2 * TILE_BYTES, qbase, ld, q0, T);
    load_tile(sQD + buf * 2 * TILE_BYTES + TILE_BYTES, dobase, inner, q0, T);
    if (threadIdx.x < TILE) {
      const int q = q0 + threadIdx.x;
      s_stat[buf * 2 * TILE + threadIdx.x] = q < T ? lse2[int64_t(bh) * T + q] : INFINITY;     // P = 0 for padded query rows
      s_stat[buf * 2 * TILE + TILE + threadIdx.x] = q < T ? dsum[int64_t(bh) * T + q] : 0.f;
    }
  };
  load_tile(sK, qbase + inner, ld, k0, T);
  load_tile(sV, qbase + 2 * inner, ld, k0, T);
  fetch_q_tile(0, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  uint32_t kf[4][4], vf[4][4];             // this warp's 16 keys as A operands, for the whole kernel
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    ldmatrix_x4(swz(sK, warp * 16 + (mi & 1) * 8 + r8, ks * 2 + (mi >> 1)), kf[ks][0], kf[ks][1], kf[ks][2], kf[ks][3]);
    ldmatrix_x4(swz(sV, warp * 16 + (mi & 1) * 8 + r8, ks * 2 + (mi >> 1)), vf[ks][0], vf[ks][1], vf[ks][2], vf[ks][3]);
  }
  float dk[8][4], dv[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
  const float sl2 = 0.125f * 1.4426950408889634f;
  const bool key_a = k0 + warp * 16 + g < T, key_b = k0 + warp * 16 + g + 8 < T;

  int buf = 0;
  for (int q0 = 0; q0 < T; q0 += TILE, buf ^= 1) {
    // the next tile streams in while this one is processed (its buffer was last read two syncs ago)
    if (q0 + TILE < T) fetch_q_tile(q0 + TILE, buf ^ 1);
    cp_async_commit();
    const uint32_t sQ = sQD + buf * 2 * TILE_BYTES, sD = sQ + TILE_BYTES;
    const float* s_lse = s_stat + buf * 2 * TILE;
    const float* s_d = s_lse + TILE;
    // S^T = K_w Q^T, P^T = exp2(S^T * sl2 - lse2[q])
    float st[8][4];
    mma_abt<kDT>(st, kf, sQ, lane);
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float la = s_lse[nt * 8 + 2 * tg], lb = s_lse[nt * 8 + 2 * tg + 1];
      st[nt][0] = key_a ? ex2_approx(fmaf(st[nt][0], sl2, -la)) : 0.f;
      st[nt][1] = key_a ? ex2_approx(fmaf(st[nt][1], sl2, -lb)) : 0.f;
      st[nt][2] = key_b ? ex2_approx(fmaf(st[nt][2], sl2, -la)) : 0.f;
      st[nt][3] = key_b ? ex2_approx(fmaf(st[nt][3], sl2, -lb)) : 0.f;
      const uint32_t va = pack2<kDT>(st[nt][0], st[nt][1]), vb = pack2<kDT>(st[nt][2], st[nt][3]);
      if ((nt & 1) == 0) { pf[nt >> 1][0] = va; pf[nt >> 1][1] = vb; }
      else               { pf[nt >> 1][2] = va; pf[nt >> 1][3] = vb; }
    }
    mma_ab<kDT>(dv, pf, sD, lane);          // dV_w += P^T dO
    // dP^T = V_w dO^T, dS^T = P^T o (dP^T - D[q]) / 8
    float dp[8][4];
    mma_abt<kDT>(dp, vf, sD, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float da = s_d[nt * 8 + 2 * tg], db = s_d[nt * 8 + 2 * tg + 1];
      const uint32_t va = pack2<kDT>(st[nt][0] * (dp[nt][0] - da) * 0.125f, st[nt][1] * (dp[nt][1] - db) * 0.125f);
      const uint32_t vb = pack2<kDT>(st[nt][2] * (dp[nt][2] - da) * 0.125f, st[nt][3] * (dp[nt][3] - db) * 0.125f);
      if ((nt & 1) == 0) { pf[nt >> 1][0] = va; pf[nt >> 1][1] = vb; }
      else               { pf[nt >> 1][2] = va; pf[nt >> 1][3] = vb; }
      sts32(swz(sT, warp * 16 + g, nt) + uint32_t(tg) * 4, va);            // staging tile [key][q] for dQ
      sts32(swz(sT, warp * 16 + g + 8, nt) + uint32_t(tg) * 4, vb);
    }
    mma_ab<kDT>(dk, pf, sQ, lane);          // dK_w += dS^T Q
    __syncthreads();                        // all 64 keys of dS^T are staged
    // dQ[16 rows of this warp] += dS K_j: A = (dS^T)^T read transposed from the staging tile
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t a[4];
      ldmatrix_x4_trans(swz(sT, ks * 16 + (mi >> 1) * 8 + r8, warp * 2 + (mi & 1)), a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(swz(sK, ks * 16 + (mi & 1) * 8 + r8, dpi * 2 + (mi >> 1)), b0, b1, b2, b3);
        mma_16816<kDT>(dq[2 * dpi], a, b0, b1);
        mma_16816<kDT>(dq[2 * dpi + 1], a, b2, b3);
      }
    }
    const int qa = q0 + warp * 16 + g, qb = qa + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = h * DH + nt * 8 + 2 * tg;
      // one 8-byte vector atomic per column pair (sm_90+)
      if (qa < T) atomicAdd(reinterpret_cast<float2*>(dq_acc + (int64_t(b) * T + qa) * inner + c), make_float2(dq[nt][0], dq[nt][1]));
      if (qb < T) atomicAdd(reinterpret_cast<float2*>(dq_acc + (int64_t(b) * T + qb) * inner + c), make_float2(dq[nt][2], dq[nt][3]));
    }
    cp_async_wait<0>();                     // the next tile has landed ...
    __syncthreads();                        // ... for everyone, and this tile's Q / dO / dS^T are no longer read
  }
  // dK, dV of this warp's 16 keys
  const int ka = k0 + warp * 16 + g, kb = ka + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = nt * 8 + 2 * tg;
    if (ka < T) {
      *reinterpret_cast<uint32_t*>(dqbase + int64_t(ka) * ld + inner + c) = pack2<kDT>(dk[nt][0], dk[nt][1]);
      *reinterpret_cast<uint32_t*>(dqbase + int64_t(ka) * ld + 2 * inner + c) = pack2<kDT>(dv[nt][0], dv[nt][1]);
    }
    if (kb < T) {
      *reinterpret_cast<uint32_t*>(dqbase + int64_t(kb) * ld + inner + c) = pack2<kDT>(dk[nt][2], dk[nt][3]);
      *reinterpret_cast<uint32_t*>(dqbase + int64_t(kb) * ld + 2 * inner + c) = pack2<kDT>(dv[nt][2], dv[nt][3]);
    }
  }
}

// dqkv[r, 0 .. inner) = cast(dq_acc[r, :])  (the q third of the [rows, 3 inner] gradient)
template <int kDT>
__global__ void __launch_bounds__(256)
dq_convert_kernel(const float* __restrict__ dq_acc, uint16_t* __restrict__ dqkv, int64_t rows, int inner) {
  const int per_row = inner >> 3;
  const int64_t total = rows * per_row;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / per_row;
    const int c = int(i - r * per_row) * 8;
    const float4 a = *reinterpret_cast<const float4*>(dq_acc + r * inner + c), b = *reinterpret_cast<const float4*>(dq_acc + r * inner + c + 4);
    uint4 o;
    o.x = pack2<kDT>(a.x, a.y); o.y = pack2<kDT>(a.z, a.w); o.z = pack2<kDT>(b.x, b.y); o.w = pack2<kDT>(b.z, b.w);
    *reinterpret_cast<uint4*>(dqkv + r * 3 * inner + c) = o;
  }
}

}  // namespace

int launch_attention_bwd_stats(cudaStream_t st, const void* qkv, const void* o_fwd, const void* d_out, float* lse2, float* dsum,
                               int batch, int T, int heads, int dtype) {
  if (batch <= 0 || T <= 0 || heads <= 0) return fail(VITB200_ERR_INVALID, "attention_bwd_stats: empty problem");
  if (int64_t(batch) * heads > 65535) return fail(VITB200_ERR_INVALID, "attention_bwd_stats: batch * heads exceeds grid.y");
  const dim3 grid(unsigned((T + TILE - 1) / TILE), unsigned(batch * heads));
  const uint16_t* q16 = static_cast<const uint16_t*>(qkv);
  const uint16_t* o16 = static_cast<const uint16_t*>(o_fwd);
  const uint16_t* d16 = static_cast<const uint16_t*>(d_out);
  if (dtype == DT_F16) attn_bwd_stats_kernel<DT_F16><<<grid, 128, 0, st>>>(q16, o16, d16, lse2, dsum, T, heads);
  else if (dtype == DT_BF16) attn_bwd_stats_kernel<DT_BF16><<<grid, 128, 0, st>>>(q16, o16, d16, lse2, dsum, T, heads);
  else return fail(VITB200_ERR_INVALID, "attention_bwd_stats: dtype must be bf16 or fp16");
  VB_LAUNCH_CHECK("attn_bwd_stats_kernel");
  return 0;
}

// fp32 dQ accumulator [rows, inner] -> the q third of dqkv (16-bit); shared with the streamed tcgen05 adjoint
int launch_dq_convert(cudaStream_t st, const float* dq_acc, void* dqkv, int64_t rows, int inner, int dtype) {
  const unsigned cgrid = unsigned(std::min<int64_t>((rows * (inner / 8) + 255) / 256, int64_t(sm_count()) * 16));
  if (dtype == DT_F16) dq_convert_kernel<DT_F16><<<cgrid, 256, 0, st>>>(dq_acc, static_cast<uint16_t*>(dqkv), rows, inner);
  else if (dtype == DT_BF16) dq_convert_kernel<DT_BF16><<<cgrid, 256, 0, st>>>(dq_acc, static_cast<uint16_t*>(dqkv), rows, inner);
  else return fail(VITB200_ERR_INVALID, "attention_bwd: dtype must be bf16 or fp16");
  VB_LAUNCH_CHECK("dq_convert_kernel");
  return 0;
}

size_t attention_bwd_flash_workspace_floats(int batch, int T, int heads) {
  // lse2 + D per (image, head, token), fp32 dQ accumulator [batch * T, heads * 64]
  return 2 * size_t(round_up(int64_t(batch) * heads * T, 64)) + size_t(batch) * T * heads * DH;
}

int launch_attention_bwd_flash(cudaStream_t st, const void* qkv, const void* o_fwd, const void* d_out, void* dqkv,
                               float* workspace, int batch, int T, int heads, int dtype) {
  if (batch <= 0 || T <= 0 || heads <= 0) return fail(VITB200_ERR_INVALID, "attention_bwd: empty problem");
  if (dtype != DT_BF16 && dtype != DT_F16) return fail(VITB200_ERR_INVALID, "attention_bwd: dtype must be bf16 or fp16");
  if (int64_t(batch) * heads > 65535) return fail(VITB200_ERR_INVALID, "attention_bwd: batch * heads exceeds grid.y; chunk the batch");
  const size_t n_stat = size_t(round_up(int64_t(batch) * heads * T, 64));   // keeps the fp32 dQ accumulator 256-byte aligned
  float* lse2 = workspace;
  float* dsum = workspace + n_stat;
  float* dq_acc = workspace + 2 * n_stat;
  const int64_t rows = int64_t(batch) * T;
  const int inner = heads * DH;
  VB_CUDA(cudaMemsetAsync(dq_acc, 0, size_t(rows) * inner * sizeof(float), st));
  const dim3 grid(unsigned((T + TILE - 1) / TILE), unsigned(batch * heads));
  const uint16_t* q16 = static_cast<const uint16_t*>(qkv);
  const uint16_t* o16 = static_cast<const uint16_t*>(o_fwd);
  const uint16_t* d16 = static_cast<const uint16_t*>(d_out);
  uint16_t* g16 = static_cast<uint16_t*>(dqkv);
  const unsigned cgrid = unsigned(std::min<int64_t>((rows * (inner / 8) + 255) / 256, int64_t(sm_count()) * 16));
  constexpr size_t kMainSmem = 7 * TILE_BYTES + 2 * 2 * TILE * sizeof(float);     // 57 KB: three CTAs per SM
  static PerDevice<bool> configured_on;   // the smem opt-in is per (function, device)
  if (bool& configured = configured_on.here(); !configured) {
    VB_CUDA(cudaFuncSetAttribute(attn_bwd_flash_kernel<DT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMainSmem)));
    VB_CUDA(cudaFuncSetAttribute(attn_bwd_flash_kernel<DT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMainSmem)));
    configured = true;
  }
  if (dtype == DT_F16) {
    attn_bwd_stats_kernel<DT_F16><<<grid, 128, 0, st>>>(q16, o16, d16, lse2, dsum, T, heads);
    VB_LAUNCH_CHECK("attn_bwd_stats_kernel");
    attn_bwd_flash_kernel<DT_F16><<<grid, 128, kMainSmem, st>>>(q16, d16, lse2, dsum, g16, dq_acc, T, heads);
    VB_LAUNCH_CHECK("attn_bwd_flash_kernel");
    dq_convert_kernel<DT_F16><<<cgrid, 256, 0, st>>>(dq_acc, g16, rows, inner);
  } else {
    attn_bwd_stats_kernel<DT_BF16><<<grid, 128, 0, st>>>(q16, o16, d16, lse2, dsum, T, heads);
    VB_LAUNCH_CHECK("attn_bwd_stats_kernel");
    attn_bwd_flash_kernel<DT_BF16><<<grid, 128, kMainSmem, st>>>(q16, d16, lse2, dsum, g16, dq_acc, T, heads);
    VB_LAUNCH_CHECK("attn_bwd_flash_kernel");
    dq_convert_kernel<DT_BF16><<<cgrid, 256, 0, st>>>(dq_acc, g16, rows, inner);
  }
  VB_LAUNCH_CHECK("dq_convert_kernel");
  return 0;
}

}  // namespace vb
