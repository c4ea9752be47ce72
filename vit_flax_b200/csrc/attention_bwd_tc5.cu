// attention_bwd_tc5.cu -- attention adjoint on tcgen05 / TMEM for sequences that fit one (image, head) on chip
// (T <= 208 tokens: every 224-px /16 config).  Adjoint of vit.py:69-79 per (image, head):
//
//   P  = exp2(Q K^T * sl2 - lse2)            sl2 = 64^-0.5 * log2(e), lse2 = log2 sum_j exp2(s_ij sl2) from the forward
//   dV = P^T dO          dP = dO V^T          D_i = sum_d dO_id O_id
//   dS = P o (dP - D) * 64^-0.5               dQ = dS K            dK = dS^T Q
//
// One work item = (image, head); persistent CTAs, 16 warps.  Q, K, V, dO of the item sit in shared memory (TMA boxes cut
// out of the to_qkv output / the cotangent, 128-byte swizzle).  The item is walked in (key tile, query tile) blocks of
// 128 x 128 (the tail tiles are 128 x 80 / 80 x 128 at T = 197); per block, all five products run on the tensor cores:
//
//   S  = Q_q K_k^T   (SS, both K-major)                      -> TMEM, fp32
//   dP = dO_q V_k^T  (SS)                                    -> TMEM, fp32
//        thread = query row (TMEM lane): p = exp2(s sl2 - lse2), ds = p (dp - D) / 8; P and dS go to shared memory as
//        16-bit tiles of 128 rows x 64-key atoms (128-byte swizzle).  ONE such tile serves both orientations:
//          as a K-major  A operand  [M = query, K = key]  (dQ += dS K_k)
//          as an MN-major A operand [M = key,   K = query] (dV += P^T dO_q, dK += dS^T Q_q) -- the transposes of
//        vit.py:73-78's adjoint are descriptor bits, nothing is moved.
//   dV_k += P^T dO_q,  dK_k += dS^T Q_q   accumulate over the query tiles in TMEM (64 columns each)
//   dQ_q += dS K_k                        accumulates over the key tiles in TMEM (64 columns per query tile)
//
// S and dP of a block are computed in two 64-key halves with their own barriers: the halves ARE the double buffer -- the
// math warps work on one half (and release it as soon as it is in registers) while the other, or the next block's, is
// computed; the MMA thread issues the next block's halves ahead of this block's dV / dK / dQ, and the dS tile is
// double-buffered, so neither side waits for the other's long phase.
// TMEM: S 128 + dP 128 + dV 64 + dK 64 + dQ 2 x 64 = 512 columns.  The softmax is NOT recomputed from scratch: the row
// log-sum-exp comes from the forward kernel (attention_tc5 writes it when asked), so no pass needs a row maximum and the
// blocks are independent.  Warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4..11 math (thread = query row; the two
// warps of a lane quarter split the block's key columns), 12..15 epilogue (dK / dV per key tile, dQ per item -> 16-bit
// rows of dqkv).  Bound: the XU pipe -- one ex2 and two fp32->16-bit packs per score, 16 cycles per score pair per SMSP.
#include <type_traits>

#include "common.h"
#include "ptx.cuh"

namespace vb {
namespace {

constexpr int DH = 64;
constexpr int BT = 128;                      // block edge (UMMA M)
constexpr int NTHR = 512;
constexpr int ROWS_MAX = 208;                // rows of Q / K / V / dO kept per item
constexpr int IN_BYTES = ROWS_MAX * 128;     // 26 KB per operand
constexpr int ATOM_BYTES = BT * 128;         // 128 rows x 64 keys of 16 bits
constexpr int OFF_Q = 0, OFF_K = IN_BYTES, OFF_V = 2 * IN_BYTES, OFF_DO = 3 * IN_BYTES;
constexpr int OFF_P = 4 * IN_BYTES, OFF_DS = OFF_P + 2 * ATOM_BYTES;
constexpr int OFF_BAR = OFF_DS + 4 * ATOM_BYTES;      // two dS tiles (consecutive blocks alternate), one P tile
constexpr int SMEM_TOTAL = OFF_BAR + 256 + 1024;
constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;   // TMEM columns

template <int kDT>
__global__ void __launch_bounds__(NTHR, 1)
attention_bwd_tc5_kernel(const __grid_constant__ CUtensorMap tmQKV,   // qkv [B, T, 3I], box 208 rows x 64
                         const __grid_constant__ CUtensorMap tmDO,    // d_out [B, T, I], box 208 rows x 64
                         const uint16_t* __restrict__ o_fwd, const float* __restrict__ lse2,
                         uint16_t* __restrict__ dqkv, int T, int heads, int items) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base + OFF_Q, sK = base + OFF_K, sV = base + OFF_V, sDO = base + OFF_DO;
  const uint32_t sP = base + OFF_P, sDS = base + OFF_DS, bars = base + OFF_BAR;
  const uint32_t in_full = bars, in_empty = bars + 8, pds_ready = bars + 16, p_free = bars + 24,
                 ds_free0 = bars + 32 /* [2] */, dkv_ready = bars + 48, dkv_free = bars + 56, dq_ready = bars + 64,
                 dq_free = bars + 72, sdp_ready0 = bars + 80 /* [2]: per 64-key half of a block */,
                 sdp_free0 = bars + 96 /* [2] */, tmem_slot = bars + 112;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gbase + OFF_BAR + 112);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int inner = heads * DH;
  const int TP = (T + 15) & ~15;                             // rows the MMAs see (zero-filled past T by TMA)
  const int ntile = (T + BT - 1) / BT;                       // key tiles == query tiles (1 or 2)
  auto ext = [&](int t) { return min(BT, TP - t * BT); };    // valid (padded) rows of tile t: 128, or 80 at T = 197

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmQKV); prefetch_tmap(&tmDO); }
  if (warp == 1 && lane == 0) {
    mbar_init(in_full, 1);   mbar_init(in_empty, 1);
    mbar_init(sdp_ready0, 1); mbar_init(sdp_ready0 + 8, 1);
    mbar_init(sdp_free0, 8);  mbar_init(sdp_free0 + 8, 8);
    mbar_init(pds_ready, 8); mbar_init(p_free, 1);
    mbar_init(ds_free0, 1);  mbar_init(ds_free0 + 8, 1);
    mbar_init(dkv_ready, 1); mbar_init(dkv_free, 4);
    mbar_init(dq_ready, 1);  mbar_init(dq_free, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer: Q, K, V, dO of one (image, head) per item =====================
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int b = item / heads, h = item - b * heads;
        mbar_wait(in_empty, uint32_t(it & 1) ^ 1u);
        mbar_arrive_expect_tx(in_full, 4 * IN_BYTES);
        tma_load_3d(sQ, &tmQKV, in_full, h * DH, 0, b);
        tma_load_3d(sK, &tmQKV, in_full, inner + h * DH, 0, b);
        tma_load_3d(sV, &tmQKV, in_full, 2 * inner + h * DH, 0, b);
        tma_load_3d(sDO, &tmDO, in_full, h * DH, 0, b);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr int fmt = kDT == DT_F16 ? 0 : 1;
      const uint32_t d_s = tmem_base + COL_S, d_dp = tmem_base + COL_DP, d_dv = tmem_base + COL_DV, d_dk = tmem_base + COL_DK;
      constexpr uint32_t idesc_t = umma_idesc_16(BT, DH, fmt, 1, 1);          // A and B MN-major: dV, dK
      constexpr uint32_t idesc_q = umma_idesc_16(BT, DH, fmt, 1, 0);          // A K-major, B MN-major: dQ
      // S = Q_q K_k^T and dP = dO_q V_k^T of block (kt, qt), in two 64-key halves with their own barriers: the halves are
      // the double buffer -- the math warps work on one while the other is computed, with no extra TMEM
      auto issue_half = [&](int kt, int qt, int hf, uint32_t bph) {
        const int nkh = max(0, min(64, ext(kt) - 64 * hf));                   // keys of this half: 64, 16 (T = 197 tail) or 0
        mbar_wait(sdp_free0 + 8u * hf, bph ^ 1u);
        if (nkh > 0) {
          const uint32_t k0 = sK + kt * ATOM_BYTES + hf * 8192, v0 = sV + kt * ATOM_BYTES + hf * 8192;
          const uint32_t q0 = sQ + qt * ATOM_BYTES, do0 = sDO + qt * ATOM_BYTES;
          const uint32_t idesc_s = umma_idesc_16(BT, nkh, fmt, 0, 0);         // [128 q] x [nkh keys], both K-major
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_bf16_ss<1>(d_s + 64 * hf, umma_desc_k_sw128(q0 + k * 32), umma_desc_k_sw128(k0 + k * 32), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_bf16_ss<1>(d_dp + 64 * hf, umma_desc_k_sw128(do0 + k * 32), umma_desc_k_sw128(v0 + k * 32), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(sdp_ready0 + 8u * hf);
      };
      int it = 0, blk = 0, kti = 0;
      const int nb = ntile * ntile;                        // blocks per item, key tile outermost
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        mbar_wait(in_full, uint32_t(it & 1));
        issue_half(0, 0, 0, uint32_t(blk & 1));
        issue_half(0, 0, 1, uint32_t(blk & 1));
        for (int j = 0; j < nb; ++j, ++blk) {
          const int kt = j / ntile, qt = j - kt * ntile;
          const int nk = ext(kt), nq = ext(qt);
          const uint32_t k0 = sK + kt * ATOM_BYTES, q0 = sQ + qt * ATOM_BYTES, do0 = sDO + qt * ATOM_BYTES;
          const uint32_t bph = uint32_t(blk & 1);
          const uint32_t ds_tile = sDS + uint32_t(blk & 1) * 2 * ATOM_BYTES;
          // the next block's first half as soon as this block's has been read, its second half right after this block's math
          if (j + 1 < nb) issue_half((j + 1) / ntile, (j + 1) % ntile, 0, bph ^ 1u);
          mbar_wait(pds_ready, bph);                       // P and dS of this block are in shared memory
          if (j + 1 < nb) issue_half((j + 1) / ntile, (j + 1) % ntile, 1, bph ^ 1u);
          // ---- dV_k += P^T dO_q : A = the P tile read MN-major (M = key, K = query), B = dO_q MN-major ----
          if (qt == 0) { mbar_wait(dkv_free, uint32_t(kti & 1) ^ 1u); }       // the previous key tile's dK / dV were drained
          tc_fence_after();
          for (int kk = 0; kk < nq / 16; ++kk)
            umma_bf16_ss<1>(d_dv, umma_desc_mn_sw128_wide(sP + kk * 2048, ATOM_BYTES), umma_desc_mn_sw128(do0 + kk * 2048),
                            idesc_t, (qt != 0 || kk != 0) ? 1u : 0u);
          umma_commit(p_free);
          // ---- dK_k += dS^T Q_q (same shape as dV) and dQ_q += dS K_k (A = the dS tile read K-major) ----
          if (j == 0) mbar_wait(dq_free, uint32_t(it & 1) ^ 1u);              // the previous item's dQ was drained
          tc_fence_after();
          for (int kk = 0; kk < nq / 16; ++kk)
            umma_bf16_ss<1>(d_dk, umma_desc_mn_sw128_wide(ds_tile + kk * 2048, ATOM_BYTES), umma_desc_mn_sw128(q0 + kk * 2048),
                            idesc_t, (qt != 0 || kk != 0) ? 1u : 0u);
          for (int kk = 0; kk < nk / 16; ++kk)
            umma_bf16_ss<1>(tmem_base + COL_DQ + qt * DH, umma_desc_k_sw128(ds_tile + (kk >> 2) * ATOM_BYTES + (kk & 3) * 32),
                            umma_desc_mn_sw128(k0 + kk * 2048), idesc_q, (kt != 0 || kk != 0) ? 1u : 0u);
          umma_commit(ds_free0 + 8u * uint32_t(blk & 1));
          if (qt == ntile - 1) { umma_commit(dkv_ready); ++kti; }
        }
        umma_commit(dq_ready);
        umma_commit(in_empty);
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 12) {
    // ===================== math: thread = query row, the two warps of a lane quarter split the key columns ==========
    const int w = warp - 4, q = w & 3, half = w >> 2;
    const int r = q * 32 + lane;                           // row inside the query tile == TMEM lane
    const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
    const float sl2 = 0.125f * 1.4426950408889634f;
    int it = 0, blk = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int b = item / heads, h = item - b * heads;
      mbar_wait(in_full, uint32_t(it & 1));
      // per row: lse2 from the forward, D = sum_d dO O (dO from the swizzled smem tile, O from global)
      float L[2], Dv[2];
#pragma unroll
      for (int qt = 0; qt < 2; ++qt) {
        const int qrow = qt * BT + r;
        L[qt] = 0.f; Dv[qt] = 0.f;
        if (qt < ntile && qrow < T) {
          L[qt] = __ldg(lse2 + (int64_t(item) * T + qrow));
          const uint4* orow = reinterpret_cast<const uint4*>(o_fwd + (int64_t(b) * T + qrow) * inner + h * DH);
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 ov = __ldg(orow + c);
            uint32_t d0, d1, d2, d3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3)
                         : "r"(sDO + uint32_t(qrow) * 128u + (uint32_t(c ^ (qrow & 7)) << 4)));
            const uint32_t dd[4] = {d0, d1, d2, d3}, oo[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc = fmaf(to_f32<kDT>(uint16_t(dd[e] & 0xFFFFu)), to_f32<kDT>(uint16_t(oo[e] & 0xFFFFu)), acc);
              acc = fmaf(to_f32<kDT>(uint16_t(dd[e] >> 16)), to_f32<kDT>(uint16_t(oo[e] >> 16)), acc);
            }
          }
          Dv[qt] = acc;
        }
      }
      for (int kt = 0; kt < ntile; ++kt) {
        const int nk = ext(kt);
        for (int qt = 0; qt < ntile; ++qt, ++blk) {
          const uint32_t bph = uint32_t(blk & 1);
          const int qrow = qt * BT + r;
          const bool row_ok = qrow < T;
          const float lse = L[qt], Dq8 = Dv[qt] * 0.125f;
          const uint32_t ds_tile = sDS + uint32_t(blk & 1) * 2 * ATOM_BYTES;
          bool tiles_free = false;                         // waited for lazily: the first exponentials overlap the MMAs
          for (int hf = 0; hf < 2; ++hf) {                 // the block's two 64-key halves (TMEM double buffer)
            const int nkh = max(0, min(64, nk - 64 * hf));
            const int wcols = nkh / 2;                     // this warp's share of the half: 32, 8 (T = 197 tail) or 0
            const int cw = 64 * hf + half * wcols;         // first of this warp's columns
            mbar_wait(sdp_ready0 + 8u * hf, bph);
            tc_fence_after();
            // the block has no key or row past T (warp-uniform): no masking selects in the inner loop
            const bool full = __all_sync(0xffffffffu, row_ok) && kt * BT + 64 * hf + nkh <= T;
            // NN (32 or 8) key columns: scores and dP out of TMEM (the half is released at once), P and dS into the smem tiles
            auto chunk = [&](auto nn_tag, int c0, bool release) {
              constexpr int NN = decltype(nn_tag)::value;
              uint32_t sv[NN], dv[NN];
              if constexpr (NN == 32) {
                tmem_ld_32x32b_x32p(t_lane + COL_S + c0, sv);
                tmem_ld_32x32b_x32p(t_lane + COL_DP + c0, dv);
              } else {
                tmem_ld_32x32b_x8(t_lane + COL_S + c0, sv);
                tmem_ld_32x32b_x8(t_lane + COL_DP + c0, dv);
              }
              tmem_ld_wait();
              if (release) {                               // the warp's last values of this half are in registers:
                tc_fence_before();                         // the next block's half may be computed
                __syncwarp();
                if (lane == 0) mbar_arrive(sdp_free0 + 8u * hf);
              }
              uint32_t pk[NN / 2], dk[NN / 2];
#pragma unroll
              for (int e = 0; e < NN; e += 2) {
                float p0 = ex2_approx(fmaf(__uint_as_float(sv[e]), sl2, -lse));
                float p1 = ex2_approx(fmaf(__uint_as_float(sv[e + 1]), sl2, -lse));
                if (!full) {                               // rows / keys past T hold arbitrary bits
                  if (!(row_ok && kt * BT + c0 + e < T)) p0 = 0.f;
                  if (!(row_ok && kt * BT + c0 + e + 1 < T)) p1 = 0.f;
                }
                float d0 = p0 * fmaf(__uint_as_float(dv[e]), 0.125f, -Dq8);          // dS = P o (dP - D) / 8
                float d1 = p1 * fmaf(__uint_as_float(dv[e + 1]), 0.125f, -Dq8);
                if (!full) {
                  if (p0 == 0.f) d0 = 0.f;
                  if (p1 == 0.f) d1 = 0.f;
                }
                pk[e >> 1] = pack2<kDT>(p0, p1);
                dk[e >> 1] = pack2<kDT>(d0, d1);
              }
              if (!tiles_free) {
                mbar_wait(p_free, bph ^ 1u);               // the previous block's dV MMAs have read the P tile
                mbar_wait(ds_free0 + 8u * uint32_t(blk & 1), (uint32_t(blk >> 1) & 1u) ^ 1u);   // dK / dQ of two blocks ago this dS tile
                tiles_free = true;
              }
#pragma unroll
              for (int g = 0; g < NN / 8; ++g) {           // 8 key columns -> one 16-byte chunk of each tile
                const int c = c0 + g * 8;
                const uint32_t off = uint32_t(c >> 6) * ATOM_BYTES + uint32_t(r) * 128u + (uint32_t(((c & 63) >> 3) ^ (r & 7)) << 4);
                st_shared_v4(sP + off, pk[g * 4 + 0], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
                st_shared_v4(ds_tile + off, dk[g * 4 + 0], dk[g * 4 + 1], dk[g * 4 + 2], dk[g * 4 + 3]);
              }
            };
            if (wcols == 32) chunk(std::integral_constant<int, 32>{}, cw, true);
            else if (wcols > 0) {                          // 8, 16 or 24 columns (short sequences, the T = 197 tail)
              for (int c = 0; c < wcols; c += 8) chunk(std::integral_constant<int, 8>{}, cw + c, c + 8 >= wcols);
            } else {                                       // the half is empty
              __syncwarp();
              if (lane == 0) mbar_arrive(sdp_free0 + 8u * hf);
            }
          }
          fence_proxy_async_smem();                        // generic-proxy tile writes -> the MMAs' async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive(pds_ready);
        }
      }
    }
  } else if (warp >= 12) {
    // ===================== epilogue: dK / dV per key tile, dQ per item -> 16-bit rows of dqkv =====================
    const int q = warp & 3, r = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(q * 32) << 16);
    const int64_t ld = 3 * int64_t(inner);
    auto store_row = [&](uint32_t col, uint16_t* dst) {    // 64 fp32 TMEM columns of this lane -> 64 16-bit values
      uint32_t v[64];
      tmem_ld_32x32b_x32p(t_lane + col, v);
      tmem_ld_32x32b_x32p(t_lane + col + 32, v + 32);
      tmem_ld_wait();
      if (dst != nullptr) {
#pragma unroll
        for (int g = 0; g < 8; ++g)
          reinterpret_cast<uint4*>(dst)[g] = make_uint4(pack2<kDT>(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1])),
                                                        pack2<kDT>(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3])),
                                                        pack2<kDT>(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5])),
                                                        pack2<kDT>(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7])));
      }
    };
    int it = 0, kti = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int b = item / heads, h = item - b * heads;
      uint16_t* rowbase = dqkv + int64_t(b) * T * ld + h * DH;
      for (int kt = 0; kt < ntile; ++kt, ++kti) {
        const int key = kt * BT + r;
        mbar_wait(dkv_ready, uint32_t(kti & 1));
        tc_fence_after();
        store_row(COL_DK, key < T ? rowbase + int64_t(key) * ld + inner : nullptr);
        store_row(COL_DV, key < T ? rowbase + int64_t(key) * ld + 2 * inner : nullptr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dkv_free);
      }
      mbar_wait(dq_ready, uint32_t(it & 1));
      tc_fence_after();
      for (int qt = 0; qt < ntile; ++qt) {
        const int qrow = qt * BT + r;
        store_row(COL_DQ + qt * DH, qrow < T ? rowbase + int64_t(qrow) * ld : nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int kDT>
int launch_t(cudaStream_t st, const void* qkv, const void* o_fwd, const void* d_out, void* dqkv, const float* lse2, int batch,
             int T, int heads) {
  static PerDevice<bool> configured_on;
  if (bool& configured = configured_on.here(); !configured) {
    VB_CUDA(cudaFuncSetAttribute(attention_bwd_tc5_kernel<kDT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    configured = true;
  }
  const int inner = heads * DH;
  CUtensorMap tq, tdo;
  int rc;
  if ((rc = make_tmap_3d_16(&tq, qkv, batch, T, 3 * inner, 3 * inner, ROWS_MAX, kDT))) return rc;
  if ((rc = make_tmap_3d_16(&tdo, d_out, batch, T, inner, inner, ROWS_MAX, kDT))) return rc;
  const int items = batch * heads;
  const int grid = items < sm_count() ? items : sm_count();
  VB_CUDA(launch_kernel(attention_bwd_tc5_kernel<kDT>, dim3(grid), dim3(NTHR), SMEM_TOTAL, st, 1, tq, tdo,
                        static_cast<const uint16_t*>(o_fwd), lse2, static_cast<uint16_t*>(dqkv), T, heads, items));
  VB_LAUNCH_CHECK("attention_bwd_tc5_kernel");
  return 0;
}

}  // namespace

bool attention_bwd_tc5_supports(int T) { return T >= 1 && T <= ROWS_MAX; }

int launch_attention_bwd_tc5(cudaStream_t st, const void* qkv, const void* o_fwd, const void* d_out, void* dqkv,
                             const float* lse2, int batch, int T, int heads, int dtype) {
  if (batch <= 0 || T <= 0 || heads <= 0) return fail(VITB200_ERR_INVALID, "attention_bwd: empty problem");
  if (!attention_bwd_tc5_supports(T)) return fail(VITB200_ERR_UNSUPPORTED, "attention_bwd_tc5: T > 208");
  if (lse2 == nullptr) return fail(VITB200_ERR_INVALID, "attention_bwd_tc5: needs the forward's row log-sum-exp");
  if (dtype == DT_BF16) return launch_t<DT_BF16>(st, qkv, o_fwd, d_out, dqkv, lse2, batch, T, heads);
  if (dtype == DT_F16) return launch_t<DT_F16>(st, qkv, o_fwd, d_out, dqkv, lse2, batch, T, heads);
  return fail(VITB200_ERR_INVALID, "attention_bwd: dtype must be bf16 or fp16");
}

}  // namespace vb
