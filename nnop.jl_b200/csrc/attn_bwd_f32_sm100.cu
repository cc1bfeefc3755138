// attn_bwd_f32_sm100.cu -- Float32 (E = 64) flash attention backward on tcgen05 tensor cores.
//
// Same algorithm and warp roles as attn_bwd_sm100.cu (one CTA per 128-key block, loop over the
// q blocks, dK/dV accumulated in TMEM, dQ reduced into global memory), for the reference's default
// element type (every reference test is Float32; README config C1).  The tensor cores have no fp32
// mode with fp32-grade products, so every operand is carried as TWO fp16 terms, x ~ hi + lo
// (22 significant bits, [hi(64) | lo(64)] side by side in a 128-wide row, written by
// attn_split_f32_rows), and every product keeps the three significant cross terms:
//    S^T  = Kh Qh^T + Kh Ql^T + Kl Qh^T                      12 MMAs of K = 16   -> TMEM [0,128)
//    dP^T = Vh dOh^T + Vh dOl^T + Vl dOh^T                   12                  -> TMEM [128,256)
//    P^T  = Ph + Pl (split in registers, both in TMEM over S^T)
//    dV' += (Ph + Pl)^T [dOh | dOl]                          16 (A in TMEM)      -> TMEM [256,384)
//    dS^T = dSh + dSl (split in registers, two smem tiles)
//    dQ'  = (dSh + dSl) [Kh | Kl]                            16                  -> aliases dP^T
//    dK' += (dSh + dSl)^T [Qh | Ql]                          16                  -> TMEM [384,512)
// X' = [X.h-part | X.l-part] accumulates both terms of the second operand side by side (N = 128); the
// consumer adds the two 64-column halves.  fp32 accumulation throughout; measured max abs error vs
// an fp64 evaluation ~1e-6 (tolerance 1e-4).  fp16's narrow range is kept out of the picture by carrying every
// tensor as x' = x * 2^-e_x (its own binary exponent; scale block, internal.h F32Mult) and folding the powers
// of two back into the logit scale, delta and the epilogue multipliers.  Shared memory is full (two dS tiles), so Q_i and dO_i are
// single-buffered and S^T(i+1) is issued after dK(i): slower per FLOP than the 16-bit kernel, still
// an order of magnitude faster than the fp32 SIMT path.  Dense layout, kpad_mask, GQA, ragged sizes.
#include "common.cuh"
#include "internal.h"

namespace nnop {
namespace {

constexpr int kThreads = 512;
constexpr float kLog2e = 1.4426950408889634f;
// dK' / dV' live in fp32 TMEM accumulators whose additions are not round-to-nearest (measured bias
// ~2^-25 of the accumulator per MMA).  Every kFlushEvery q blocks (16 MMAs each) the drain warpgroup
// moves them out with exact fp32 reduce-adds into dk / dv and the accumulation restarts from zero.
constexpr int kFlushEvery = 4;
using T = __half;

struct Params {
  const float* lse2p;   // (B*QH, QLp) lse * log2e, +inf padded
  const float* deltap;  // (B*QH, QLp)
  int QL, KL, QH, KH, QLp, causal;
  float scale, scale_log2;
  const float* mult;    // scale block multipliers (internal.h F32Mult): the operands are q', k', v', dO' = x * 2^-e_x
  const uint8_t* kpad;
  // additive bias (BIAS kernel): head-major fp32 copies (B, QH, QL, KLp), see attn_pair.cu
  const float* pair_t;
  float* dpair_t;
  int KLp;
};

struct Smem {
  static constexpr int kTile = 128 * 128 * 2;  // 32 KB: [hi | lo] tile of 128 rows
  static constexpr int kBox = 128 * 64 * 2;    // 16 KB: one 64-column box (box 0 = hi, box 1 = lo)
  static constexpr int kK = 0;
  static constexpr int kV = kK + kTile;
  static constexpr int kQ = kV + kTile;
  static constexpr int kdO = kQ + kTile;
  static constexpr int kdSh = kdO + kTile;     // dS^T hi: 128 keys x 128 q, two boxes
  static constexpr int kdSl = kdSh + kTile;    // dS^T lo
  static constexpr int kStage = kdSl + kTile;  // 128 rows x 64 fp32 (two boxes of 32 floats)
  static constexpr int kStat = kStage + kTile; // lse2[128], delta[128]
  static constexpr int kBar = kStat + 1024;
  static constexpr int kNumBars = 14;
  static constexpr int kTotal = kBar + kNumBars * 8 + 16;
};

__device__ __forceinline__ float bf_lo(uint32_t packed) { return unpack_lo<T>(packed); }
__device__ __forceinline__ float bf_hi(uint32_t packed) { return unpack_hi<T>(packed); }

// BIAS = true: S^T gets pair^T added before the exponential and dpair = dS is written out, both through
// the head-major copies with lanes along the key axis (coalesced), as in attn_bwd_sm100.cu.
template <bool BIAS>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_f32_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                    const __grid_constant__ CUtensorMap tm_dk, const __grid_constant__ CUtensorMap tm_dv,
                    const __grid_constant__ CUtensorMap tm_dq, const Params p) {
  using S = Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + S::kK;
  uint8_t* sV = smem + S::kV;
  uint8_t* sQ = smem + S::kQ;
  uint8_t* sdO = smem + S::kdO;
  uint8_t* sdSh = smem + S::kdSh;
  uint8_t* sdSl = smem + S::kdSl;
  uint8_t* sStage = smem + S::kStage;
  float* s_lse = reinterpret_cast<float*>(smem + S::kStat);
  float* s_del = s_lse + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = bars + 2;
  uint64_t* do_full = bars + 3;
  uint64_t* do_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* dp_full = bars + 7;
  uint64_t* ds_full = bars + 8;
  uint64_t* dq_full = bars + 9;
  uint64_t* dq_empty = bars + 10;
  uint64_t* dkdv_full = bars + 11;
  uint64_t* flush_full = bars + 12;   // MMA -> drain: accumulators of the last kFlushEvery steps complete
  uint64_t* flush_done = bars + 13;   // drain -> MMA: accumulators read out, may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + S::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x;
  const int k0 = j * 128;
  const int hk = blockIdx.y, b = blockIdx.z;
  const int QL = p.QL, KL = p.KL;
  const int g = p.QH / p.KH;
  const int bh_kv = b * p.KH + hk;
  const int nq = (QL + 127) >> 7;
  const int i0 = p.causal ? j : 0;
  const int nqi = nq > i0 ? nq - i0 : 0;
  bool key_keep = true;
  if (p.kpad) {
    const int kr = k0 + (threadIdx.x & 127);
    key_keep = kr < KL && p.kpad[static_cast<int64_t>(b) * p.KL + kr] != 0;
  }
  const bool any_key = p.kpad ? __syncthreads_or(key_keep) != 0 : true;
  const int n_it = any_key ? nqi * g : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dq);
    tma_prefetch_desc(&tm_dk);
    tma_prefetch_desc(&tm_dv);
    mbar_init(kv_full, 1);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(do_full, 1);
    mbar_init(do_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);
    mbar_init(dp_full, 1);
    mbar_init(ds_full, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);
    mbar_init(dkdv_full, 1);
    mbar_init(flush_full, 1);
    mbar_init(flush_done, 4);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 384;

  if (warp < 4) {
    setmaxnreg_dec<88>();
    if (warp == 0 && lane == 0 && n_it > 0) {
      // ================================ TMA producer =================================
      mbar_arrive_expect_tx(kv_full, 2 * S::kTile);
#pragma unroll
      for (int bx = 0; bx < 2; ++bx) {
        tma_load_3d(sK + bx * S::kBox, &tm_k, kv_full, bx * 64, k0, bh_kv);
        tma_load_3d(sV + bx * S::kBox, &tm_v, kv_full, bx * 64, k0, bh_kv);
      }
      for (int it = 0; it < n_it; ++it) {
        const int bh_q = b * p.QH + hk * g + it / nqi;
        const int q0 = (i0 + it % nqi) * 128;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, S::kTile + 1024);
#pragma unroll
        for (int bx = 0; bx < 2; ++bx) tma_load_3d(sQ + bx * S::kBox, &tm_q, q_full, bx * 64, q0, bh_q);
        const int64_t soff = static_cast<int64_t>(bh_q) * p.QLp + q0;
        bulk_load_1d(s_lse, p.lse2p + soff, 512, q_full);
        bulk_load_1d(s_del, p.deltap + soff, 512, q_full);
        mbar_wait(do_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(do_full, S::kTile);
#pragma unroll
        for (int bx = 0; bx < 2; ++bx) tma_load_3d(sdO + bx * S::kBox, &tm_do, do_full, bx * 64, q0, bh_q);
      }
    } else if (warp == 1 && n_it > 0) {
      // ================================ MMA issuer ===================================
      constexpr uint32_t id_kk = make_idesc_f16(128, 128, false, false, false);  // S^T, dP^T (fp16 operands)
      constexpr uint32_t id_tv = make_idesc_f16(128, 128, false, false, true);   // dV', dK' (B MN-major)
      constexpr uint32_t id_mm = make_idesc_f16(128, 128, false, true, true);    // dQ'
      const uint32_t tm = uniform_u32(tmem_base);
      const uint32_t sbase = uniform_u32(smem_u32(smem));
      const uint64_t kmaj = make_smem_desc_sw128(sbase, 16, 1024);
      const uint64_t mnmaj = make_smem_desc_sw128(sbase, S::kBox, 1024);
      const uint32_t k_lo = desc_lo(kmaj), k_hi = desc_hi(kmaj);
      const uint32_t m_lo = desc_lo(mnmaj), m_hi = desc_hi(mnmaj);
      // D[128 x 128] = Ah Bh^T + Ah Bl^T + Al Bh^T for K-major [hi | lo] tiles at byte offsets a0 / b0
      auto mma_3term = [&](uint32_t dcol, uint32_t a0, uint32_t b0) {
        if (elect_one()) {
#pragma unroll
          for (int part = 0; part < 3; ++part)
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint32_t aoff = (part == 2 ? S::kBox : 0) + k4 * 32;
              const uint32_t boff = (part == 1 ? S::kBox : 0) + k4 * 32;
              umma_ss_lo(tm + dcol, k_lo, (a0 + aoff) >> 4, k_hi, k_lo, (b0 + boff) >> 4, k_hi, id_kk,
                         (part | k4) ? 1u : 0u);
            }
        }
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) tc_commit(bar);
      };
      mbar_wait(kv_full, 0);
      mbar_wait(q_full, 0);
      tc_fence_after();
      mma_3term(kColS, S::kK, S::kQ);
      commit(s_full);
      mbar_wait(do_full, 0);
      tc_fence_after();
      mma_3term(kColDP, S::kV, S::kdO);
      commit(dp_full);
      int nflush = 0;
      for (int it = 0; it < n_it; ++it) {
        const bool fresh = it % kFlushEvery == 0;   // first step after a flush (or the first at all)
        const uint32_t acc = fresh ? 0u : 1u;
        if (fresh && it > 0) {                      // the drain warpgroup must have read dV' / dK' out
          mbar_wait(flush_done, (nflush - 1) & 1);
        }
        // dV' += (Ph + Pl)^T [dOh | dOl]: q columns 16*ks.. of P^T live in the half (ks >> 2) of S^T
        mbar_wait(p_full, it & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t a = tm + kColS + (ks >> 2) * 64 + (ks & 3) * 8;
            umma_ts_lo(tm + kColDV, a, m_lo, (S::kdO + ks * 2048) >> 4, m_hi, id_tv, (acc | (ks > 0)) ? 1u : 0u);
            umma_ts_lo(tm + kColDV, a + 32, m_lo, (S::kdO + ks * 2048) >> 4, m_hi, id_tv, 1u);
          }
        }
        commit(do_empty);
        mbar_wait(ds_full, it & 1);
        tc_fence_after();
        // dK' first: Q' is single-buffered here, so its reload (and with it S^T(i+1)) can only start when
        // dK' has read it -- issued before dQ', the load runs under the dQ' MMAs instead of after them
        // dK' += (dSh + dSl)^T [Qh | Ql]   (A: dS^T K-major; B: Q' MN-major, N = 128)
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t off = (ks >> 2) * S::kBox + (ks & 3) * 32;
            umma_ss_lo(tm + kColDK, k_lo, (S::kdSh + off) >> 4, k_hi, m_lo, (S::kQ + ks * 2048) >> 4, m_hi, id_tv,
                       (acc | (ks > 0)) ? 1u : 0u);
            umma_ss_lo(tm + kColDK, k_lo, (S::kdSl + off) >> 4, k_hi, m_lo, (S::kQ + ks * 2048) >> 4, m_hi, id_tv, 1u);
          }
        }
        commit(q_empty);
        // dQ' = (dSh + dSl) [Kh | Kl]   (A: dS^T tiles viewed MN-major; B: K' MN-major, N = 128)
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            umma_ss_lo(tm + kColDP, m_lo, (S::kdSh + ks * 2048) >> 4, m_hi, m_lo, (S::kK + ks * 2048) >> 4, m_hi,
                       id_mm, ks > 0 ? 1u : 0u);
            umma_ss_lo(tm + kColDP, m_lo, (S::kdSl + ks * 2048) >> 4, m_hi, m_lo, (S::kK + ks * 2048) >> 4, m_hi,
                       id_mm, 1u);
          }
        }
        commit(dq_full);
        if ((it + 1) % kFlushEvery == 0 && it + 1 < n_it) {
          commit(flush_full);
          ++nflush;
        }
        if (it + 1 < n_it) {
          mbar_wait(q_full, (it + 1) & 1);
          tc_fence_after();
          mma_3term(kColS, S::kK, S::kQ);
          commit(s_full);
          mbar_wait(do_full, (it + 1) & 1);
          mbar_wait(dq_empty, it & 1);   // dP^T's columns hold dQ'_i until the drain warpgroup read them
          tc_fence_after();
          mma_3term(kColDP, S::kV, S::kdO);
          commit(dp_full);
        }
      }
      commit(dkdv_full);
    }
  } else if (warp < 12) {
    // ================================ compute warpgroups ===============================
    setmaxnreg_inc<136>();
    const int half = (warp - 4) >> 2;  // which 64 q columns
    const int wq = warp & 3;
    const int row = wq * 32 + lane;    // key row within the block
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const int c0 = half * 64;
    const float sl2 = p.scale_log2 * __ldg(p.mult + F32Mult::kLogits);   // S = 2^(e_q + e_k) * q' k'^T
    const float dpair_mul = BIAS ? __ldg(p.mult + F32Mult::kDPair) : 1.f;
    for (int it = 0; it < n_it; ++it) {
      const int i = i0 + it % nqi;
      float pf[64];
      int64_t boff = 0;       // (BIAS) element offset of this thread's first bias / dpair entry
      int nqv = 0;            // (BIAS) valid q columns of this half
      if constexpr (BIAS) {
        // bias of this step, parked in pf (dead here) while S^T is still being computed
        const int bh_q = b * p.QH + hk * g + it / nqi;
        const int qb = i * 128 + c0;
        boff = (static_cast<int64_t>(bh_q) * QL + qb) * p.KLp + k0 + row;
        nqv = min(64, QL - qb);
        // unconditional loads (64 in flight per thread): out-of-range rows / keys re-read a valid
        // entry instead of being predicated -- their P is zeroed or never used further down
        const float* bb = p.pair_t + static_cast<int64_t>(bh_q) * QL * p.KLp + min(k0 + row, p.KLp - 1);
        if (nqv == 64) {
          const float* bp = bb + static_cast<int64_t>(qb) * p.KLp;
#pragma unroll
          for (int c = 0; c < 64; ++c) pf[c] = bp[static_cast<int64_t>(c) * p.KLp];
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) pf[c] = bb[static_cast<int64_t>(min(qb + c, QL - 1)) * p.KLp];
        }
      }
      mbar_wait(q_full, it & 1);   // lse2 / delta of this q block have landed
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      uint32_t sr[2][32];
      tmem_ld_x32(tmem_base + lane_off + kColS + c0, sr[0]);
      tmem_ld_x32(tmem_base + lane_off + kColS + c0 + 32, sr[1]);
      tmem_ld_wait();
      const float4* l4 = reinterpret_cast<const float4*>(s_lse + c0);
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float4 l = l4[u];
        // with a bias: pf holds pair^T; (pair * log2e - lse2) replaces -lse2
        const float b0 = BIAS ? fmaf(pf[4 * u + 0], kLog2e, -l.x) : -l.x;
        const float b1 = BIAS ? fmaf(pf[4 * u + 1], kLog2e, -l.y) : -l.y;
        const float b2 = BIAS ? fmaf(pf[4 * u + 2], kLog2e, -l.z) : -l.z;
        const float b3 = BIAS ? fmaf(pf[4 * u + 3], kLog2e, -l.w) : -l.w;
        pf[4 * u + 0] = fast_exp2(fmaf(__uint_as_float(sr[u >> 3][(4 * u + 0) & 31]), sl2, b0));
        pf[4 * u + 1] = fast_exp2(fmaf(__uint_as_float(sr[u >> 3][(4 * u + 1) & 31]), sl2, b1));
        pf[4 * u + 2] = fast_exp2(fmaf(__uint_as_float(sr[u >> 3][(4 * u + 2) & 31]), sl2, b2));
        pf[4 * u + 3] = fast_exp2(fmaf(__uint_as_float(sr[u >> 3][(4 * u + 3) & 31]), sl2, b3));
      }
      if (p.causal && i == j) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (row > c0 + c) pf[c] = 0.f;
      }
      if (!key_keep) {
#pragma unroll
        for (int c = 0; c < 64; ++c) pf[c] = 0.f;
      }
      {
        // P^T = Ph + Pl, both packed into this warpgroup's OWN 64 columns of S^T: [Ph 32 | Pl 32]
        uint32_t ph[32], pl[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          ph[c] = pack2<T>(pf[2 * c], pf[2 * c + 1]);
          pl[c] = pack2<T>(pf[2 * c] - bf_lo(ph[c]), pf[2 * c + 1] - bf_hi(ph[c]));
        }
        tmem_st_x32(tmem_base + lane_off + kColS + c0, ph);
        tmem_st_x32(tmem_base + lane_off + kColS + c0 + 32, pl);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- dS^T = P^T o (dP^T - delta), split into two 16-bit tiles ----
      mbar_wait(dp_full, it & 1);
      tc_fence_after();
      tmem_ld_x32(tmem_base + lane_off + kColDP + c0, sr[0]);
      tmem_ld_x32(tmem_base + lane_off + kColDP + c0 + 32, sr[1]);
      tmem_ld_wait();
      const float4* d4 = reinterpret_cast<const float4*>(s_del + c0);
      uint8_t* drow_h = sdSh + half * S::kBox + row * 128;
      uint8_t* drow_l = sdSl + half * S::kBox + row * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const float4 da = d4[2 * ch], db = d4[2 * ch + 1];
        const float dl[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
        float ds[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = 8 * ch + e;
          ds[e] = pf[c] * (__uint_as_float(sr[c >> 5][c & 31]) - dl[e]);
        }
        uint4 vh, vl;
        vh.x = pack2<T>(ds[0], ds[1]); vl.x = pack2<T>(ds[0] - bf_lo(vh.x), ds[1] - bf_hi(vh.x));
        vh.y = pack2<T>(ds[2], ds[3]); vl.y = pack2<T>(ds[2] - bf_lo(vh.y), ds[3] - bf_hi(vh.y));
        vh.z = pack2<T>(ds[4], ds[5]); vl.z = pack2<T>(ds[4] - bf_lo(vh.z), ds[5] - bf_hi(vh.z));
        vh.w = pack2<T>(ds[6], ds[7]); vl.w = pack2<T>(ds[6] - bf_lo(vh.w), ds[7] - bf_hi(vh.w));
        *reinterpret_cast<uint4*>(drow_h + ((ch ^ (row & 7)) << 4)) = vh;
        *reinterpret_cast<uint4*>(drow_l + ((ch ^ (row & 7)) << 4)) = vl;
        if constexpr (BIAS) {  // dpair = dS (before the 1/sqrt(E) that dQ / dK carry)
          if (k0 + row < KL) {
            float* dp = p.dpair_t + boff;
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (8 * ch + e < nqv) dp[static_cast<int64_t>(8 * ch + e) * p.KLp] = ds[e] * dpair_mul;
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    // ---- epilogue: dV (half 0) / dK (half 1): add the two 64-column halves, fp32 TMA store ----
    if (n_it > 0) {
      mbar_wait(dkdv_full, 0);
      tc_fence_after();
    }
    {
      const uint32_t tsrc = tmem_base + lane_off + (half ? kColDK : kColDV);
      const float mul = half ? p.scale * __ldg(p.mult + F32Mult::kDK) : __ldg(p.mult + F32Mult::kDV);
      uint8_t* stage = half ? sdO : sQ;  // 32 KB each, free by now
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t rh[32], rl[32];
        if (n_it > 0) {
          tmem_ld_x32(tsrc + c * 32, rh);
          tmem_ld_x32(tsrc + 64 + c * 32, rl);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int x = 0; x < 32; ++x) rh[x] = rl[x] = 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float4 v;
          v.x = (__uint_as_float(rh[4 * u + 0]) + __uint_as_float(rl[4 * u + 0])) * mul;
          v.y = (__uint_as_float(rh[4 * u + 1]) + __uint_as_float(rl[4 * u + 1])) * mul;
          v.z = (__uint_as_float(rh[4 * u + 2]) + __uint_as_float(rl[4 * u + 2])) * mul;
          v.w = (__uint_as_float(rh[4 * u + 3]) + __uint_as_float(rl[4 * u + 3])) * mul;
          *reinterpret_cast<float4*>(stage + c * S::kBox + row * 128 + ((u ^ (row & 7)) << 4)) = v;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + half, 128);
      if (wq == 0 && lane == 0 && n_it > 0) {   // dk / dv were zeroed by the launcher: add the remainder
#pragma unroll
        for (int bx = 0; bx < 2; ++bx)
          tma_reduce_add_3d(half ? &tm_dk : &tm_dv, stage + bx * S::kBox, bx * 32, k0, bh_kv);
        bulk_commit();
        bulk_wait_read<0>();
      }
    }
  } else {
    // ================================ dQ drain warpgroup ===============================
    setmaxnreg_inc<152>();
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const bool issuer = (warp == 12 && lane == 0);
    const float dq_mul = p.scale * __ldg(p.mult + F32Mult::kDQ);
    const float dk_mul = p.scale * __ldg(p.mult + F32Mult::kDK), dv_mul = __ldg(p.mult + F32Mult::kDV);
    int nflush = 0;
    for (int it = 0; it < n_it; ++it) {
      const int bh_q = b * p.QH + hk * g + it / nqi;
      const int q0 = (i0 + it % nqi) * 128;
      mbar_wait(dq_full, it & 1);
      tc_fence_after();
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(tmem_base + lane_off + kColDP + c * 32, r[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      if (issuer) bulk_wait_read<0>();   // the previous step's reduce has finished reading the stage
      named_bar_sync(3, 128);
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float4 v;   // dQ = scale * 2^(e_dO + e_v + e_k) * (dQ'[:, 0:64] + dQ'[:, 64:128])
          v.x = (__uint_as_float(r[c][4 * u + 0]) + __uint_as_float(r[c + 2][4 * u + 0])) * dq_mul;
          v.y = (__uint_as_float(r[c][4 * u + 1]) + __uint_as_float(r[c + 2][4 * u + 1])) * dq_mul;
          v.z = (__uint_as_float(r[c][4 * u + 2]) + __uint_as_float(r[c + 2][4 * u + 2])) * dq_mul;
          v.w = (__uint_as_float(r[c][4 * u + 3]) + __uint_as_float(r[c + 2][4 * u + 3])) * dq_mul;
          *reinterpret_cast<float4*>(sStage + c * S::kBox + row * 128 + ((u ^ (row & 7)) << 4)) = v;
        }
      fence_proxy_async_smem();
      named_bar_sync(3, 128);
      if (issuer) {
        tma_reduce_add_3d(&tm_dq, sStage, 0, q0, bh_q);
        tma_reduce_add_3d(&tm_dq, sStage + S::kBox, 32, q0, bh_q);
        bulk_commit();
      }
      if ((it + 1) % kFlushEvery == 0 && it + 1 < n_it) {
        // ---- flush dV' and dK' (see kFlushEvery) ----
        mbar_wait(flush_full, (nflush++) & 1);
        tc_fence_after();
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          const uint32_t tsrc = tmem_base + lane_off + (which ? kColDK : kColDV);
          const float mul = which ? dk_mul : dv_mul;
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld_x32(tsrc + c * 32, r[c]);
          tmem_ld_wait();
          if (which == 1) {   // both accumulators are in registers / staged: the MMA warp may go on
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(flush_done);
          }
          if (issuer) bulk_wait_read<0>();
          named_bar_sync(3, 128);
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              float4 v;
              v.x = (__uint_as_float(r[c][4 * u + 0]) + __uint_as_float(r[c + 2][4 * u + 0])) * mul;
              v.y = (__uint_as_float(r[c][4 * u + 1]) + __uint_as_float(r[c + 2][4 * u + 1])) * mul;
              v.z = (__uint_as_float(r[c][4 * u + 2]) + __uint_as_float(r[c + 2][4 * u + 2])) * mul;
              v.w = (__uint_as_float(r[c][4 * u + 3]) + __uint_as_float(r[c + 2][4 * u + 3])) * mul;
              *reinterpret_cast<float4*>(sStage + c * S::kBox + row * 128 + ((u ^ (row & 7)) << 4)) = v;
            }
          fence_proxy_async_smem();
          named_bar_sync(3, 128);
          if (issuer) {
            tma_reduce_add_3d(which ? &tm_dk : &tm_dv, sStage, 0, k0, bh_kv);
            tma_reduce_add_3d(which ? &tm_dk : &tm_dv, sStage + S::kBox, 32, k0, bh_kv);
            bulk_commit();
          }
        }
      }
    }
    if (issuer) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// prep: delta' = 2^-(e_dO + e_v) * rowsum(dO o O) over the 64 fp32 columns, lse2 (padded, +inf => P = 0), dq := 0
__global__ void __launch_bounds__(256)
attn_bwd_f32_prep_kernel(float* __restrict__ deltap, float* __restrict__ lse2p, float* __restrict__ dq,
                         const float* __restrict__ dO, const float* __restrict__ o,
                         const float* __restrict__ lse, int QL, int QLp, int64_t n_rows_p,
                         const float* __restrict__ mult, int E) {
  const int lpr = E >> 2;         // lanes (one float4 each) per E-float row: 16, 8 or 4
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const int64_t rowp = gid / lpr;
  const int li = static_cast<int>(gid % lpr);
  if (rowp >= n_rows_p) return;
  const int64_t bh = rowp / QLp;
  const int q = static_cast<int>(rowp % QLp);
  float acc = 0.f;
  if (q < QL) {
    const int64_t off = (bh * QL + q) * E + li * 4;
    const float4 a = *reinterpret_cast<const float4*>(dO + off);
    const float4 c = *reinterpret_cast<const float4*>(o + off);
    acc = a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
    *reinterpret_cast<float4*>(dq + off) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int sft = 1; sft < lpr; sft <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
  if (li == 0) {
    const float l = q < QL ? lse[bh * QL + q] : INFINITY;
    deltap[rowp] = q < QL ? acc * __ldg(mult + F32Mult::kDeltaInv) : 0.f;   // delta' pairs with dP' = dO' v'^T
    lse2p[rowp] = l == -INFINITY ? INFINITY : l * kLog2e;
  }
}

inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

}  // namespace

size_t attn_f32_bwd_workspace_bytes(int QL, int KL, int QH, int KH, int B) {
  const size_t QLp = static_cast<size_t>((QL + 127) / 128) * 128;
  const size_t BH = static_cast<size_t>(B) * QH, BHk = static_cast<size_t>(B) * KH;
  return 2 * align256(BH * QLp * sizeof(float)) + (2 * BH * QL + 2 * BHk * KL) * 128 * 2 + kF32ScaleBytes;
}

int attn_f32_bwd(const AttnParams& a) {
  using S = Smem;
  const int QLp = ((a.QL + 127) / 128) * 128;
  const int64_t BH = static_cast<int64_t>(a.B) * a.QH, BHk = static_cast<int64_t>(a.B) * a.KH;
  char* ws = reinterpret_cast<char*>(a.delta);
  const size_t stat_bytes = align256(static_cast<size_t>(BH) * QLp * sizeof(float));
  float* deltap = reinterpret_cast<float*>(ws);
  float* lse2p = reinterpret_cast<float*>(ws + stat_bytes);
  T* qs = reinterpret_cast<T*>(ws + 2 * stat_bytes);
  T* dos = qs + BH * a.QL * 128;
  T* ks = dos + BH * a.QL * 128;
  T* vs = ks + BHk * a.KL * 128;
  void* blk = vs + BHk * a.KL * 128;   // scale block: |x|max, exponents, multipliers (internal.h)
  const int E = a.E;   // 16 / 32 / 64: narrower rows are zero-padded into the same [hi(64) | lo(64)] operand rows
  if (int rc = attn_f32_scales(blk, a.q, BH * a.QL * E, a.k, BHk * a.KL * E, a.v, BHk * a.KL * E, a.dO,
                               BH * a.QL * E, a.stream))
    return rc;
  {
    const int64_t n_rows_p = BH * QLp;
    const int64_t threads = n_rows_p * (a.E / 4);
    attn_bwd_f32_prep_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, a.stream>>>(
        deltap, lse2p, static_cast<float*>(a.dq), static_cast<const float*>(a.dO),
        static_cast<const float*>(a.o), a.lse, a.QL, QLp, n_rows_p, f32_mults(blk), a.E);
    NNOP_LAUNCH_CHECK();
  }
  {
    // q, dO, k, v -> [hi | lo] fp16 rows and the zeroing of the dk / dv accumulators: one launch (six 5-10 us nodes
    // before; at the reference's README shape they were ~6 % of the backward)
    F32StageJobs jobs;
    jobs.out[0] = qs; jobs.in[0] = a.q; jobs.rows[0] = BH * a.QL; jobs.scale[0] = f32_in_scale(blk, 0);
    jobs.out[1] = dos; jobs.in[1] = a.dO; jobs.rows[1] = BH * a.QL; jobs.scale[1] = f32_in_scale(blk, 3);
    jobs.out[2] = ks; jobs.in[2] = a.k; jobs.rows[2] = BHk * a.KL; jobs.scale[2] = f32_in_scale(blk, 1);
    jobs.out[3] = vs; jobs.in[3] = a.v; jobs.rows[3] = BHk * a.KL; jobs.scale[3] = f32_in_scale(blk, 2);
    jobs.zero[0] = a.dk; jobs.zero_floats[0] = BHk * a.KL * E;
    jobs.zero[1] = a.dv; jobs.zero_floats[1] = BHk * a.KL * E;
    if (int rc = attn_stage_f32(jobs, E, a.stream)) return rc;
  }
  alignas(64) CUtensorMap tq, tk, tv, tdo, tdk, tdv, tdq;
  const uint64_t bhq = static_cast<uint64_t>(BH), bhk = static_cast<uint64_t>(BHk);
  if (int rc = make_tmap_3d(&tq, qs, NNOP_F16, 128, a.QL, bhq, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tdo, dos, NNOP_F16, 128, a.QL, bhq, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tk, ks, NNOP_F16, 128, a.KL, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tv, vs, NNOP_F16, 128, a.KL, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tdk, a.dk, NNOP_F32, E, a.KL, bhk, 32, 128)) return rc;
  if (int rc = make_tmap_3d(&tdv, a.dv, NNOP_F32, E, a.KL, bhk, 32, 128)) return rc;
  if (int rc = make_tmap_3d(&tdq, a.dq, NNOP_F32, E, a.QL, bhq, 32, 128)) return rc;
  const bool bias = a.pair != nullptr;
  auto kern = bias ? attn_bwd_f32_kernel<true> : attn_bwd_f32_kernel<false>;
  NNOP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
  if (bias && !a.pair_t_ready)
    if (int rc = attn_pair_to_head_major(a)) return rc;
  Params bp;
  bp.lse2p = lse2p; bp.deltap = deltap;
  bp.QL = a.QL; bp.KL = a.KL; bp.QH = a.QH; bp.KH = a.KH; bp.QLp = QLp; bp.causal = a.causal;
  bp.scale = a.scale; bp.scale_log2 = a.scale * kLog2e; bp.mult = f32_mults(blk);
  bp.kpad = a.kpad;
  bp.pair_t = static_cast<const float*>(a.pair_t); bp.dpair_t = static_cast<float*>(a.dpair_t); bp.KLp = a.KLp;
  dim3 grid((a.KL + 127) / 128, a.KH, a.B);
  timing_begin(1, a.stream);
  kern<<<grid, kThreads, S::kTotal, a.stream>>>(tq, tk, tv, tdo, tdk, tdv, tdq, bp);
  timing_end(1, a.stream);
  NNOP_LAUNCH_CHECK();
  if (bias)
    if (int rc = attn_dpair_from_head_major(a)) return rc;
  return NNOP_OK;
}

}  // namespace nnop
