// attention_tc5.cu -- K4 (second generation): fused attention on tcgen05 / TMEM for sequences
// that fit one key block (T <= 208 tokens: every 224-px /16 config, i.e. ViT-B/16 and ViT-L/16).
//
//   out = softmax(Q K^T * 64^-0.5) V   per (image, head)                        vit.py:69-79
//
// One work item = (image, head, 128-query tile).  Persistent CTAs, 12 warps:
//   warp 0  lane 0 : TMA producer: Q tile, K, V boxes cut straight out of the to_qkv output viewed
//                    as [B, T, 3I] (the split / head rearranges of vit.py:69-71 are coordinates);
//                    rows >= T are out of bounds of the image and arrive as zeros.
//   warp 1  lane 0 : tcgen05.mma issuer:  S = Q K^T  (M=128, N=KP, K=64; SS operands)
//                                         O = P V    (M=128, N=64,  K=KP; P from TMEM, V MN-major)
//   warp 2         : TMEM allocator: two 256-column slots; S occupies [0,KP), P (16-bit, two keys
//                    per column) overwrites [0,KP/2), O lives at [128,192) of the same slot.
//   warps 4..11    : softmax + epilogue, thread = (row, column half).  The half-row of S is read
//                    ONCE into registers; row max / row sum are exchanged between the two halves
//                    through shared memory; P goes back to TMEM with tcgen05.st; the epilogue of
//                    item i-1 (O / rowsum -> 16-bit -> smem -> TMA store, rows >= T clipped) runs
//                    between the max exchange and the exponentials of item i, which frees the
//                    TMEM slot early enough for S(i+1) to be computed under softmax(i).
// The [B,h,T,T] score tensor of vit.py:73-75 never reaches HBM.
#include "common.h"
#include "ptx.cuh"

namespace vb {

namespace {

constexpr int DH = 64;
constexpr int QT = 128;                 // queries per work item (UMMA M)
constexpr int NT = 384;                 // threads per CTA
constexpr int Q_BYTES = QT * 128;       // 16 KB
constexpr int O_BYTES = QT * 128;       // 16 KB staging for the TMA store

template <int OFF, int REM>
__device__ __forceinline__ void ld_row(uint32_t taddr, uint32_t* r) {
  if constexpr (REM >= 32) {
    tmem_ld_32x32b_x32p(taddr + OFF, r + OFF);
    ld_row<OFF + 32, REM - 32>(taddr, r);
  } else if constexpr (REM >= 16) {
    tmem_ld_32x32b_x16(taddr + OFF, r + OFF);
    ld_row<OFF + 16, REM - 16>(taddr, r);
  } else if constexpr (REM >= 8) {
    tmem_ld_32x32b_x8(taddr + OFF, r + OFF);
    ld_row<OFF + 8, REM - 8>(taddr, r);
  }
}
template <int OFF, int REM>
__device__ __forceinline__ void st_row(uint32_t taddr, const uint32_t* r) {
  if constexpr (REM >= 16) {
    tmem_st_32x32b_x16(taddr + OFF, r + OFF);
    st_row<OFF + 16, REM - 16>(taddr, r);
  } else if constexpr (REM >= 8) {
    tmem_st_32x32b_x8(taddr + OFF, r + OFF);
    st_row<OFF + 8, REM - 8>(taddr, r);
  } else if constexpr (REM >= 4) {
    tmem_st_32x32b_x4(taddr + OFF, r + OFF);
    st_row<OFF + 4, REM - 4>(taddr, r);
  }
}

template <int KP>
struct Smem {
  static constexpr int KV_BYTES = KP * 128;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + 2 * Q_BYTES;
  static constexpr int OFF_V = OFF_K + 2 * KV_BYTES;
  static constexpr int OFF_O = OFF_V + 2 * KV_BYTES;
  static constexpr int OFF_BAR = OFF_O + O_BYTES;          // 16 mbarriers + tmem slot
  static constexpr int OFF_XMAX = OFF_BAR + 256;               // float [2 slots][2][128]
  static constexpr int OFF_XSUM = OFF_XMAX + 2 * 2 * 128 * 4;  // float [2 slots][2][128]
  static constexpr int TOTAL = OFF_XSUM + 2 * 2 * 128 * 4 + 1024 /*align slack*/;
  static_assert(KV_BYTES % 1024 == 0, "K/V stage must keep 1024-byte alignment");
};

template <int kDT, int KP>
__global__ void __launch_bounds__(NT, 1)
attention_tc5_kernel(const __grid_constant__ CUtensorMap tmQ,    // qkv [B,T,3I], box 128 rows
                     const __grid_constant__ CUtensorMap tmKV,   // qkv [B,T,3I], box KP rows
                     const __grid_constant__ CUtensorMap tmO,    // out [B,T,I],  box 128 rows
                     int T, int heads, int nqt, int items) {
  using L = Smem<KP>;
  constexpr int CH = KP / 2;            // S columns per thread (one half of the row)
  constexpr int PH = CH / 2;            // packed P columns per thread
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + L::OFF_Q, sK = base + L::OFF_K, sV = base + L::OFF_V, sO = base + L::OFF_O;
  const uint32_t bars = base + L::OFF_BAR;
  auto qk_full = [&](int s) { return bars + 8u * (0 + s); };
  auto v_full = [&](int s) { return bars + 8u * (2 + s); };
  auto qk_empty = [&](int s) { return bars + 8u * (4 + s); };
  auto v_empty = [&](int s) { return bars + 8u * (6 + s); };
  auto s_ready = [&](int s) { return bars + 8u * (8 + s); };
  auto p_ready = [&](int s) { return bars + 8u * (10 + s); };
  auto o_ready = [&](int s) { return bars + 8u * (12 + s); };
  auto slot_free = [&](int s) { return bars + 8u * (14 + s); };
  const uint32_t tmem_slot = bars + 8u * 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + L::OFF_BAR + 8 * 16);
  float* xmax = reinterpret_cast<float*>(gbase + L::OFF_XMAX);
  float* xsum = reinterpret_cast<float*>(gbase + L::OFF_XSUM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int inner = heads * DH;
  const int64_t first = int64_t(blockIdx.x) * items / gridDim.x;
  const int64_t last = int64_t(blockIdx.x + 1) * items / gridDim.x;
  const int n = int(last - first);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmKV);
    prefetch_tmap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(qk_full(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(qk_empty(s), 1);
      mbar_init(v_empty(s), 1);
      mbar_init(s_ready(s), 1);
      mbar_init(p_ready(s), 8);
      mbar_init(o_ready(s), 1);
      mbar_init(slot_free(s), 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < n; ++i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        const int64_t item = first + i;
        const int bh = int(item / nqt), qt = int(item - int64_t(bh) * nqt);
        const int b = bh / heads, h = bh - b * heads;
        mbar_wait(qk_empty(s), ph ^ 1u);
        mbar_arrive_expect_tx(qk_full(s), Q_BYTES + L::KV_BYTES);
        tma_load_3d(sQ + s * Q_BYTES, &tmQ, qk_full(s), h * DH, qt * QT, b);
        tma_load_3d(sK + s * L::KV_BYTES, &tmKV, qk_full(s), inner + h * DH, 0, b);
        mbar_wait(v_empty(s), ph ^ 1u);
        mbar_arrive_expect_tx(v_full(s), L::KV_BYTES);
        tma_load_3d(sV + s * L::KV_BYTES, &tmKV, v_full(s), 2 * inner + h * DH, 0, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr int fmt = kDT == DT_F16 ? 0 : 1;
      constexpr uint32_t idesc_s = umma_idesc_16(QT, KP, fmt, 0);   // B = K, K-major
      constexpr uint32_t idesc_o = umma_idesc_16(QT, DH, fmt, 1);   // B = V, MN-major
      auto issue_s = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait(slot_free(s), ph ^ 1u);
        mbar_wait(qk_full(s), ph);
        tc_fence_after();
        const uint32_t q0 = sQ + s * Q_BYTES, k0 = sK + s * L::KV_BYTES;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_bf16_ss<1>(tmem_base + s * 256, umma_desc_k_sw128(q0 + k * 32),
                          umma_desc_k_sw128(k0 + k * 32), idesc_s, k != 0 ? 1u : 0u);
        umma_commit(qk_empty(s));
        umma_commit(s_ready(s));
      };
      auto issue_pv = [&](int i) {
        const int s = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        mbar_wait(p_ready(s), ph);
        mbar_wait(v_full(s), ph);
        tc_fence_after();
        const uint32_t v0 = sV + s * L::KV_BYTES;
#pragma unroll
        for (int kk = 0; kk < KP / 16; ++kk)
          umma_bf16_ts(tmem_base + s * 256 + 128, tmem_base + s * 256 + kk * 8,
                       umma_desc_mn_sw128(v0 + kk * 2048), idesc_o, kk != 0 ? 1u : 0u);
        umma_commit(v_empty(s));
        umma_commit(o_ready(s));
      };
      if (n > 0) issue_s(0);
      for (int i = 0; i < n; ++i) {
        if (i + 1 < n) issue_s(i + 1);
        issue_pv(i);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== softmax + epilogue =====================
    const int w = warp - 4;
    const int q = w & 3;                  // TMEM lane quarter (== warp % 4)
    const int hf = w >> 2;                // column half of the score row
    const int row = q * 32 + lane;        // query row in the tile == TMEM lane
    const bool leader = (w == 0 && lane == 0);
    const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
    const float sl2 = 0.125f * 1.4426950408889634f;   // dim_head^-0.5 * log2(e)   (vit.py:66)

    // epilogue of item j: O / rowsum -> 16-bit -> swizzled smem -> TMA store.  Called by all 256
    // threads right after a group barrier (which also orders the xsum / staging-buffer hazards).
    auto epilogue = [&](int j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int64_t item = first + j;
      const int bh = int(item / nqt), qt = int(item - int64_t(bh) * nqt);
      const int b = bh / heads, h = bh - b * heads;
      mbar_wait(o_ready(s), ph);
      tc_fence_after();
      const float inv = 1.0f / (xsum[(s * 2 + 0) * 128 + row] + xsum[(s * 2 + 1) * 128 + row]);
#pragma unroll
      for (int c = 0; c < 2; ++c) {        // 2 x 16 columns of this thread's 32
        uint32_t o[16];
        tmem_ld_32x32b_x16(t_lane + uint32_t(s * 256 + 128 + hf * 32 + c * 16), o);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g) {      // 8 columns -> one 16-byte chunk
          const int chunk = hf * 4 + c * 2 + g;
          st_shared_v4(sO + uint32_t(row) * 128u + (uint32_t(chunk ^ (row & 7)) << 4),
                       pack2<kDT>(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv),
                       pack2<kDT>(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv),
                       pack2<kDT>(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv),
                       pack2<kDT>(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(4, 256);
      if (lane == 0) mbar_arrive(slot_free(s));          // TMEM slot (S/P/O of item j) reusable
      if (leader) {
        tma_store_3d(&tmO, sO, h * DH, qt * QT, b);      // rows >= T are clipped by the tensor map
        tma_store_commit();
      }
    };

    for (int i = 0; i < n; ++i) {
      const int s = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      mbar_wait(s_ready(s), ph);
      tc_fence_after();
      uint32_t sr[CH];
      ld_row<0, CH>(t_lane + uint32_t(s * 256 + hf * CH), sr);
      tmem_ld_wait();
      // half-row max.  Keys >= T come from zero-filled K rows (score 0, not -inf): they are masked,
      // but only in the 8-column groups that reach past T (warp-uniform test), so the common path
      // is one 3-input max per two scores and nothing else.
      const int nvalid = T - hf * CH;       // valid columns of this half (may be <= 0 or >= CH)
      float mx = -INFINITY;
#pragma unroll
      for (int c8 = 0; c8 < CH / 8; ++c8) {
        if ((c8 + 1) * 8 > nvalid) {
#pragma unroll
          for (int j = c8 * 8; j < c8 * 8 + 8; ++j)
            if (j >= nvalid) sr[j] = __float_as_uint(-INFINITY);
        }
#pragma unroll
        for (int j = c8 * 8; j < c8 * 8 + 8; j += 2)
          mx = fmaxf(mx, fmaxf(__uint_as_float(sr[j]), __uint_as_float(sr[j + 1])));
      }
      xmax[(s * 2 + hf) * 128 + row] = mx;       // slot-indexed: no WAR race with a slow partner
      if (leader) tma_store_wait_read<0>();              // staging buffer of epilogue(i-2) drained
      tc_fence_before();
      named_bar_sync(3, 256);                            // all S reads done; maxima + staging visible
      mx = fmaxf(mx, xmax[(s * 2 + (hf ^ 1)) * 128 + row]);
      if (i > 0) epilogue(i - 1);
      // p = exp2((s - max) * scale * log2e); un-normalised, row sum kept in fp32
      const float mneg = -mx * sl2;
      const unsigned long long sl2x2 = pack_f32x2(sl2, sl2), mnegx2 = pack_f32x2(mneg, mneg);
      unsigned long long lx2 = pack_f32x2(0.f, 0.f);
      uint32_t pp[PH];
#pragma unroll
      for (int j = 0; j < PH; ++j) {
        float a0, a1;
        unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sr[2 * j]), __uint_as_float(sr[2 * j + 1])),
                               sl2x2, mnegx2), a0, a1);
        const float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
        lx2 = add_f32x2(lx2, pack_f32x2(p0, p1));
        pp[j] = pack2<kDT>(p0, p1);
      }
      float l0, l1;
      unpack_f32x2(lx2, l0, l1);
      const float l = l0 + l1;
      xsum[(s * 2 + hf) * 128 + row] = l;
      st_row<0, PH>(t_lane + uint32_t(s * 256 + hf * PH), pp);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(s));
    }
    if (n > 0) {
      if (leader) tma_store_wait_read<0>();
      named_bar_sync(3, 256);
      epilogue(n - 1);
      if (leader) tma_store_wait<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int kDT, int KP>
int launch_kp(cudaStream_t stream, const void* qkv, void* out, int batch, int T, int heads) {
  using L = Smem<KP>;
  static bool configured = false;
  if (!configured) {
    VB_CUDA(cudaFuncSetAttribute(attention_tc5_kernel<kDT, KP>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int inner = heads * DH;
  CUtensorMap tq, tkv, to;
  int rc;
  if ((rc = make_tmap_3d_16(&tq, qkv, batch, T, 3 * inner, 3 * inner, QT, kDT))) return rc;
  if ((rc = make_tmap_3d_16(&tkv, qkv, batch, T, 3 * inner, 3 * inner, KP, kDT))) return rc;
  if ((rc = make_tmap_3d_16(&to, out, batch, T, inner, inner, QT, kDT))) return rc;
  const int nqt = ceil_div(T, QT);
  const int64_t items64 = int64_t(batch) * heads * nqt;
  if (items64 > 0x7fffffff) return fail(VITB200_ERR_INVALID, "attention: too many work items");
  const int items = int(items64);
  const int grid = items < sm_count() ? items : sm_count();
  attention_tc5_kernel<kDT, KP><<<grid, NT, L::TOTAL, stream>>>(tq, tkv, to, T, heads, nqt, items);
  VB_LAUNCH_CHECK("attention_tc5_kernel");
  return 0;
}

template <int kDT>
int launch_dt(cudaStream_t stream, const void* qkv, void* out, int batch, int T, int heads) {
  if (T <= 64) return launch_kp<kDT, 64>(stream, qkv, out, batch, T, heads);
  if (T <= 128) return launch_kp<kDT, 128>(stream, qkv, out, batch, T, heads);
  return launch_kp<kDT, 208>(stream, qkv, out, batch, T, heads);
}

}  // namespace

bool attention_tc5_supports(int T) { return T >= 1 && T <= 208; }

int launch_attention_tc5(cudaStream_t stream, const void* qkv, void* out, int batch, int T, int heads,
                         int dtype) {
  if (batch <= 0 || T <= 0 || heads <= 0) return fail(VITB200_ERR_INVALID, "attention: empty problem");
  if (!attention_tc5_supports(T)) return fail(VITB200_ERR_UNSUPPORTED, "attention_tc5: T > 208");
  if (dtype == DT_BF16) return launch_dt<DT_BF16>(stream, qkv, out, batch, T, heads);
  if (dtype == DT_F16) return launch_dt<DT_F16>(stream, qkv, out, batch, T, heads);
  return fail(VITB200_ERR_INVALID, "attention: dtype must be bf16 or fp16");
}

}  // namespace vb
