// attn_bwd_sm100.cu -- flash attention backward on tcgen05 / TMEM / TMA (bf16 / fp16, E in
// {64,128}, causal or not, GQA, ragged QL/KL).  Replaces `_flash_attention_bwd_preprocess!`
// and `_flash_attention_bwd!` (src/attention_bwd.jl:163-197, :1-161), whose grid is only
// (QH, B) workgroups of SIMT 32x32 tiles with global read-modify-write of dQ/dK/dV.
//
// Three launches:
//   1. prep:  delta = rowsum(dO o O), lse2 = lse*log2(e) (padded to 128-row blocks; padding rows
//             get lse2 = +inf so their P is exactly 0), and zeroes the fp32 dQ accumulator.
//   2. main:  one CTA per (128-key block j, kv head, batch); loops over the q heads of the GQA
//             group (dK/dV are summed inside the CTA -- no atomics, unlike :100,139) and over the
//             128-row q blocks i (causal: i >= j only).  Per step, five 128x128x128 MMAs:
//                S^T  = K_j Q_i^T          (smem x smem)            -> TMEM [0,128)
//                dP^T = V_j dO_i^T         (smem x smem)            -> TMEM [128,256)
//                dV  += P^T dO_i           (A = P^T bf16 in TMEM, aliasing S^T: each compute warpgroup packs its
//                                           64 q columns into the first 32 columns of its OWN half of S^T, so
//                                           neither can overwrite logits the other has not read yet)
//                dQ_i = dS K_j             (A = dS^T smem read MN-major) -> TMEM [128,128+E), aliasing dP^T
//                dK  += dS^T Q_i           (A = dS^T smem read K-major)
//             P^T = exp2(S^T*scale*log2e - lse2[q]) and dS^T = P^T o (dP^T - delta[q]) are computed
//             by two warpgroups (thread <-> key row, half of the q columns each).  A fourth
//             warpgroup drains dQ_i from TMEM and adds it into the fp32 accumulator with TMA
//             reduce-add (cp.reduce.async.bulk.tensor), 32 columns at a time.
//             Issue order per step: dV(i), S^T(i+1), dQ(i), dK(i), dP^T(i+1) (persistent kernel: dK(i)
//             before dQ(i)) so the tensor pipe works on step i's products while the warpgroups
//             exponentiate step i+1.
//   3. post:  dQ = T(scale * dQ_accum).
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "internal.h"

namespace nnop {
namespace {

constexpr int kBwdThreads = 512;
constexpr float kLog2e = 1.4426950408889634f;

// 0 (default) = one CTA per kv block, 1 = CTA pairs (cta_group::2) where eligible; env NNOP_BWD_PAIR /
// nnop_set_bwd_pair_mode.  The pair kernel is exact but measured slower (5 600 vs 4 000 clk per step
// on config C2): what it saves on the shared-memory port it loses to cross-CTA barrier round trips
// on the dP -> dS -> dQ -> drain -> dP chain.  Kept as an experiment; see DESIGN.md 4.2.
std::atomic<int> g_bwd_pair_mode{-1};
inline int bwd_pair_mode() {
  int m = g_bwd_pair_mode.load();
  if (m < 0) {
    const char* e = getenv("NNOP_BWD_PAIR");
    m = e ? atoi(e) : 0;
    g_bwd_pair_mode.store(m);
  }
  return m;
}

#ifdef NNOP_BWD_TRACE
// development aid: pipeline timeline of CTA (0,0,0), 16 clock64 stamps per q-block step
__device__ long long g_bwd_trace[256 * 16];
#define BWD_STAMP(i, k) do { if (tr) g_bwd_trace[(i) * 16 + (k)] = clock64(); } while (0)
// every CTA that runs on SM 0 logs (j, n_it, 6 clock stamps): per-CTA fixed costs and gaps
__device__ long long g_bwd_cta_log[1024 * 16];
__device__ int g_bwd_cta_n;
#define BWD_CTA(k) do { if (cta_slot >= 0) g_bwd_cta_log[cta_slot * 16 + (k)] = clock64(); } while (0)
#define PT_STAMP(k) do { if (blockIdx.x == 0 && lane == 0 && tl < 1024) g_bwd_cta_log[tl * 16 + (k)] = clock64(); } while (0)
#else
#define PT_STAMP(k) do { } while (0)
#define BWD_CTA(k) do { } while (0)
#define BWD_STAMP(i, k) do { } while (0)
#endif

struct BwdParams {
  const float* lse2p;   // (B*QH, QLp) lse * log2e, +inf padded
  const float* deltap;  // (B*QH, QLp)
  int QL, KL, QH, KH, QLp, causal;
  float scale, scale_log2;
  // packed variable-length mode (cu_q != nullptr): sequence z = blockIdx.z; QL / KL are maxima;
  // the padded statistics of sequence z start at row (cu_q[z] / 128 + z) * 128 of a QLp-long axis
  const int* cu_q;
  const int* cu_k;
  void* dk_ptr;
  void* dv_ptr;
  int64_t total_k;
  const uint8_t* kpad;  // (B, KL) key padding mask, 1 = attend, or nullptr (dense mode only)
  // persistent kernel, packed mode: tile_pre[z] = number of (kv block, kv head) tiles of the
  // sequences before z (nseq + 1 entries, built by packed_tile_prefix_kernel)
  const int* tile_pre;
  int nseq;
  // persistent kernel, dense causal mode: (batch, kv head) units are walked in groups of `lpt_group`
  // whose Q / dO streams fit in L2 together, each group heaviest kv block first (0 = unit by unit)
  int lpt_group;
  // additive bias (dense mode, BIAS kernels): head-major copies (B, QH, QL, KLp), see attn_pair.cu
  const void* pair_t;
  void* dpair_t;
  int KLp;
};

// first padded-statistics row of packed sequence z (each sequence is padded to whole 128-row blocks)
#ifdef NNOP_BWD_NO_EXP   // (timing experiments: knock out the MUFU work)
#define fast_exp2(x) (x)
#endif
__device__ __forceinline__ int packed_stat_row(int cu, int z) { return ((cu >> 7) + z) << 7; }

template <int D>
struct BwdSmem {
  static constexpr int kTile = 128 * D * 2;  // K_j, V_j, Q_i, dO_i tiles
  static constexpr int kBox = 128 * 64 * 2;  // 64-column box
  static constexpr int kNBox = D / 64;
  static constexpr int kK = 0;
  static constexpr int kV = kK + kTile;
  static constexpr int kQ = kV + kTile;         // 2 stages
  static constexpr int kdO = kQ + 2 * kTile;    // 1 stage
  static constexpr int kdS = kdO + kTile;       // 128 x 128 16-bit, two boxes
  static constexpr int kdQs = kdS + 2 * kBox;   // 2 x (128 rows x 32 fp32)
  static constexpr int kStat = kdQs + 2 * 16384;  // lse2[2][128], delta[2][128]
  static constexpr int kBar = kStat + 2048;
  static constexpr int kNumBars = 32;  // 14 used by the plain kernel, 26 (+ 32 bytes of tile ring) by the persistent one
  static constexpr int kTotal = kBar + kNumBars * 8 + 16;
};

// BIAS = true: S^T gets pair^T added before the exponential and dpair = dS (src/attention_bwd.jl:120-131)
// is written out, both through the head-major copies with lanes along the key axis (coalesced).
template <typename T, int D, bool BIAS = false>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_q,
                      const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v,
                      const __grid_constant__ CUtensorMap tm_do,
                      const __grid_constant__ CUtensorMap tm_dk,
                      const __grid_constant__ CUtensorMap tm_dv,
                      const __grid_constant__ CUtensorMap tm_dqa, const BwdParams p) {
  using S = BwdSmem<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + S::kK;
  uint8_t* sV = smem + S::kV;
  uint8_t* sQ = smem + S::kQ;
  uint8_t* sdO = smem + S::kdO;
  uint8_t* sdS = smem + S::kdS;
  uint8_t* sdQ = smem + S::kdQs;
  float* s_lse = reinterpret_cast<float*>(smem + S::kStat);        // [2][128]
  float* s_del = reinterpret_cast<float*>(smem + S::kStat + 1024);  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint64_t* kv_full = bars + 0;
  uint64_t* q_full = bars + 1;    // [2]
  uint64_t* q_empty = bars + 3;   // [2]
  uint64_t* do_full = bars + 5;
  uint64_t* do_empty = bars + 6;
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* dp_full = bars + 9;
  uint64_t* ds_full = bars + 10;
  uint64_t* dq_full = bars + 11;
  uint64_t* dq_empty = bars + 12;
  uint64_t* dkdv_full = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + S::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef NNOP_BWD_TRACE
  const bool tr = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
  int& s_cta_slot = *reinterpret_cast<int*>(smem + S::kBar + S::kNumBars * 8 + 8);  // spare bytes after the TMEM slot
  if (threadIdx.x == 0) {
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    s_cta_slot = smid == 0 ? atomicAdd(&g_bwd_cta_n, 1) : -1;
    if (s_cta_slot >= 1024) s_cta_slot = -1;
    if (s_cta_slot >= 0) g_bwd_cta_log[s_cta_slot * 16 + 0] = clock64();
  }
  __syncthreads();
  const int cta_slot = lane == 0 ? s_cta_slot : -1;
#endif

  // ---- work assignment ----------------------------------------------------------------
  const int j = blockIdx.x;  // kv block
  const int k0 = j * 128;
  const int hk = blockIdx.y;
  const bool packed = p.cu_q != nullptr;
  int QL = p.QL, KL = p.KL, q_off = 0, k_off = 0, st_off = 0, b = blockIdx.z;
  if (packed) {
    q_off = p.cu_q[blockIdx.z];
    QL = p.cu_q[blockIdx.z + 1] - q_off;
    k_off = p.cu_k[blockIdx.z];
    KL = p.cu_k[blockIdx.z + 1] - k_off;
    st_off = packed_stat_row(q_off, blockIdx.z);
    b = 0;
    if (k0 >= KL) return;  // grid sized for the longest sequence
  }
  const int g = p.QH / p.KH;
  const int bh_kv = b * p.KH + hk;
  const int nq = (QL + 127) >> 7;
  const int i0 = p.causal ? j : 0;
  const int nqi = nq > i0 ? nq - i0 : 0;
  // key padding mask (src/attention.jl:73-79): a block with no attended key does no work at all;
  // otherwise the masked key rows (thread <-> key row in the compute warpgroups) are zeroed in P^T
  bool key_keep = true;
  if (p.kpad) {
    const int kr = k0 + (threadIdx.x & 127);
    key_keep = kr < KL && p.kpad[static_cast<int64_t>(b) * p.KL + kr] != 0;
  }
  const bool any_key = p.kpad ? __syncthreads_or(key_keep) != 0 : true;
  const int n_it = any_key ? nqi * g : 0;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("nnop: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dqa);
    mbar_init(kv_full, 1);
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    mbar_init(&q_empty[0], 1);
    mbar_init(&q_empty[1], 1);
    mbar_init(do_full, 1);
    mbar_init(do_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 8);    // one arrival per compute warp
    mbar_init(dp_full, 1);
    mbar_init(ds_full, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 4);  // one arrival per drain warp
    mbar_init(dkdv_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 256 + D;
  if (threadIdx.x == 0) BWD_CTA(1);

  if (warp < 4) {
    setmaxnreg_dec<88>();  // 512 thr x 128 regs at launch -> 88 / 136 / 152 (sum = 64K regs)
    if (warp == 0 && lane == 0 && n_it > 0) {
      // ================================ TMA producer =================================
      mbar_arrive_expect_tx(kv_full, 2 * S::kTile);
#pragma unroll
      for (int bx = 0; bx < S::kNBox; ++bx) {
        tma_load_3d(sK + bx * S::kBox, &tm_k, kv_full, bx * 64, k_off + k0, bh_kv);
        tma_load_3d(sV + bx * S::kBox, &tm_v, kv_full, bx * 64, k_off + k0, bh_kv);
      }
      auto load_q = [&](int it) {
        const int s = it & 1;
        const int bh_q = b * p.QH + hk * g + it / nqi;
        const int q0 = (i0 + it % nqi) * 128;
        mbar_wait(&q_empty[s], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[s], S::kTile + 1024);
#pragma unroll
        for (int bx = 0; bx < S::kNBox; ++bx)
          tma_load_3d(sQ + s * S::kTile + bx * S::kBox, &tm_q, &q_full[s], bx * 64, q_off + q0, bh_q);
        const int64_t soff = static_cast<int64_t>(bh_q) * p.QLp + st_off + q0;
        bulk_load_1d(s_lse + s * 128, p.lse2p + soff, 512, &q_full[s]);
        bulk_load_1d(s_del + s * 128, p.deltap + soff, 512, &q_full[s]);
      };
      auto load_do = [&](int it) {
        const int bh_q = b * p.QH + hk * g + it / nqi;
        const int q0 = (i0 + it % nqi) * 128;
        mbar_wait(do_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(do_full, S::kTile);
#pragma unroll
        for (int bx = 0; bx < S::kNBox; ++bx)
          tma_load_3d(sdO + bx * S::kBox, &tm_do, do_full, bx * 64, q_off + q0, bh_q);
      };
      load_q(0);
      load_do(0);
      if (n_it > 1) load_q(1);
      for (int it = 1; it < n_it; ++it) {
        load_do(it);
        if (it + 1 < n_it) load_q(it + 1);
      }
    } else if (warp == 1 && n_it > 0) {
      // ================================ MMA issuer ===================================
      // Whole warp runs the uniform control flow and the waits; one elected lane issues.  All
      // operands are warp-uniform so descriptors stay in uniform registers (see attn_fwd_sm100.cu).
      constexpr bool BF = is_bf16<T>::value;
      constexpr uint32_t id_kk = make_idesc_f16(128, 128, BF, false, false);  // S^T, dP^T
      constexpr uint32_t id_tv = make_idesc_f16(128, D, BF, false, true);     // dV (A in TMEM), dK
      constexpr uint32_t id_mm = make_idesc_f16(128, D, BF, true, true);      // dQ
      const uint32_t tm = uniform_u32(tmem_base);
      const uint32_t sbase = uniform_u32(smem_u32(smem));
      // K-major views (LBO unused, SBO 1024) and MN-major views (LBO = next 64-element box)
      // descriptor templates: operand = template low word + (byte offset >> 4), constant high word
      const uint64_t kmaj = make_smem_desc_sw128(sbase, 16, 1024);        // K-major (SBO 1024)
      const uint64_t mnmaj = make_smem_desc_sw128(sbase, S::kBox, 1024);  // MN-major (LBO = next box)
      const uint32_t k_lo = desc_lo(kmaj), k_hi = desc_hi(kmaj);
      const uint32_t m_lo = desc_lo(mnmaj), m_hi = desc_hi(mnmaj);
      // D[128 x 128] = A[128 x D] B[128 x D]^T, both K-major tiles at byte offsets a0 / b0
      auto mma_kk = [&](uint32_t dcol, uint32_t a0, uint32_t b0) {
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t off = (ks >> 2) * S::kBox + (ks & 3) * 32;
            umma_ss_lo(tm + dcol, k_lo, (a0 + off) >> 4, k_hi, k_lo, (b0 + off) >> 4, k_hi, id_kk,
                       ks > 0 ? 1u : 0u);
          }
        }
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) tc_commit(bar);
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      mma_kk(kColS, S::kK, S::kQ);
      commit(s_full);
      mbar_wait(do_full, 0);
      tc_fence_after();
      mma_kk(kColDP, S::kV, S::kdO);
      commit(dp_full);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const uint32_t acc = it > 0 ? 1u : 0u;
        const uint32_t qoff = static_cast<uint32_t>(s * S::kTile);
        // dV += P^T dO_i
        mbar_wait(p_full, it & 1);
        tc_fence_after();
        BWD_STAMP(it, 0);
        if (it == 0) BWD_CTA(2);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ts_lo(tm + kColDV, tm + kColS + (ks >> 2) * 64 + (ks & 3) * 8, m_lo,
                       (S::kdO + ks * 2048) >> 4, m_hi, id_tv, (acc | (ks > 0)) ? 1u : 0u);
        }
        commit(do_empty);
        // S^T(i+1)
        if (it + 1 < n_it) {
          mbar_wait(&q_full[s ^ 1], ((it + 1) >> 1) & 1);
          tc_fence_after();
          mma_kk(kColS, S::kK, S::kQ + static_cast<uint32_t>((s ^ 1) * S::kTile));
          commit(s_full);
        }
        mbar_wait(ds_full, it & 1);
        tc_fence_after();
        BWD_STAMP(it, 1);
        // dK += dS^T Q_i  (A: dS^T K-major; B: Q_i MN-major) -- before dQ_i: it releases the Q stage,
        // whose reload is on the S^T(i+2) critical path (see the persistent kernel)
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t off = (ks >> 2) * S::kBox + (ks & 3) * 32;
            umma_ss_lo(tm + kColDK, k_lo, (S::kdS + off) >> 4, k_hi, m_lo, (S::kQ + qoff + ks * 2048) >> 4,
                       m_hi, id_tv, (acc | (ks > 0)) ? 1u : 0u);
          }
        }
        commit(&q_empty[s]);
        // dQ_i = dS K_j   (A: dS^T smem viewed MN-major; B: K_j MN-major)
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ss_lo(tm + kColDP, m_lo, (S::kdS + ks * 2048) >> 4, m_hi, m_lo, (S::kK + ks * 2048) >> 4,
                       m_hi, id_mm, ks > 0 ? 1u : 0u);
        }
        commit(dq_full);
        // dP^T(i+1) -- its TMEM columns hold dQ_i until the drain warpgroup has read them
        if (it + 1 < n_it) {
          mbar_wait(do_full, (it + 1) & 1);
          BWD_STAMP(it, 2);
          mbar_wait(dq_empty, it & 1);
          tc_fence_after();
          BWD_STAMP(it, 3);
          mma_kk(kColDP, S::kV, S::kdO);
          commit(dp_full);
        }
      }
      commit(dkdv_full);
      BWD_CTA(3);
    }
  } else if (warp < 12) {
    // ================================ compute warpgroups ===============================
    setmaxnreg_inc<136>();
    const int half = (warp - 4) >> 2;  // which 64 q columns
    const int wq = warp & 3;
    const int row = wq * 32 + lane;    // key row within the block
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const int c0 = half * 64;
    const float sl2 = p.scale_log2;
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      const int i = i0 + it % nqi;
      float pf[64];
      int64_t boff = 0;       // (BIAS) element offset of this thread's first bias / dpair entry
      int nqv = 0;            // (BIAS) valid q columns of this half
      if constexpr (BIAS) {
        // bias of this step, parked in pf (dead here) while S^T is still being computed:
        // pair_t[bh_q][q][k] with lanes along k -> one 128-byte line per warp and q column
        const int bh_q = b * p.QH + hk * g + it / nqi;
        const int qb = i * 128 + c0;
        boff = (static_cast<int64_t>(bh_q) * QL + qb) * p.KLp + k0 + row;
        nqv = min(64, QL - qb);
        // unconditional loads (64 in flight per thread): out-of-range rows / keys re-read a valid
        // entry instead of being predicated -- their P is zeroed or never used further down
        const T* bb = static_cast<const T*>(p.pair_t) + static_cast<int64_t>(bh_q) * QL * p.KLp +
                      min(k0 + row, p.KLp - 1);
        if (nqv == 64) {
          const T* bp = bb + static_cast<int64_t>(qb) * p.KLp;
#pragma unroll
          for (int c = 0; c < 64; ++c) pf[c] = to_f32<T>(bp[static_cast<int64_t>(c) * p.KLp]);
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) pf[c] = to_f32<T>(bb[static_cast<int64_t>(min(qb + c, QL - 1)) * p.KLp]);
        }
      }
      // ---- P^T ----
      mbar_wait(&q_full[s], (it >> 1) & 1);  // lse2 / delta of this stage have landed
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      if (wq == 0) BWD_STAMP(it, 4 + 4 * half);
      uint32_t sr[2][32];
      tmem_ld_x32(tmem_base + lane_off + kColS + c0, sr[0]);
      tmem_ld_x32(tmem_base + lane_off + kColS + c0 + 32, sr[1]);
      tmem_ld_wait();
      const float4* l4 = reinterpret_cast<const float4*>(s_lse + s * 128 + c0);
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
      if (p.causal && i == j) {  // diagonal block: key k0+row is visible to query q0+c iff row <= c
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (row > c0 + c) pf[c] = 0.f;
      }
      if ((packed && k0 + row >= KL) || !key_keep) {  // next packed sequence's key / padded key
#pragma unroll
        for (int c = 0; c < 64; ++c) pf[c] = 0.f;
      }
      {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) pk[c] = pack2<T>(pf[2 * c], pf[2 * c + 1]);
        tmem_st_x32(tmem_base + lane_off + kColS + half * 64, pk);  // inside this half's own S^T columns
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (wq == 0) BWD_STAMP(it, 5 + 4 * half);
      // ---- dS^T ----
      mbar_wait(dp_full, it & 1);
      tc_fence_after();
      if (wq == 0) BWD_STAMP(it, 6 + 4 * half);
      tmem_ld_x32(tmem_base + lane_off + kColDP + c0, sr[0]);
      tmem_ld_x32(tmem_base + lane_off + kColDP + c0 + 32, sr[1]);
      tmem_ld_wait();
      const float4* d4 = reinterpret_cast<const float4*>(s_del + s * 128 + c0);
      uint8_t* drow = sdS + half * S::kBox + row * 128;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {  // 8 columns = one 16-byte chunk
        const float4 da = d4[2 * ch], db = d4[2 * ch + 1];
        const float dl[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
        float ds[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = 8 * ch + e;
          ds[e] = pf[c] * (__uint_as_float(sr[c >> 5][c & 31]) - dl[e]);
        }
        uint4 v;
        v.x = pack2<T>(ds[0], ds[1]);
        v.y = pack2<T>(ds[2], ds[3]);
        v.z = pack2<T>(ds[4], ds[5]);
        v.w = pack2<T>(ds[6], ds[7]);
        *reinterpret_cast<uint4*>(drow + ((ch ^ (row & 7)) << 4)) = v;
        if constexpr (BIAS) {  // dpair = dS (before the 1/sqrt(E) that dQ / dK carry)
          if (k0 + row < KL) {
            T* dp = static_cast<T*>(p.dpair_t) + boff;
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (8 * ch + e < nqv) dp[static_cast<int64_t>(8 * ch + e) * p.KLp] = from_f32<T>(ds[e]);
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
      if (wq == 0) BWD_STAMP(it, 7 + 4 * half);
    }
    // ---- epilogue: dV (half 0) / dK (half 1) -> 16-bit -> swizzled smem -> TMA store ------
    if (n_it > 0) {
      mbar_wait(dkdv_full, 0);
      tc_fence_after();
    }
    if (warp == 4) BWD_CTA(4);
    {
      const uint32_t tsrc = tmem_base + lane_off + (half ? kColDK : kColDV);
      const float mul = half ? p.scale : 1.f;
      uint8_t* stage = sQ + half * S::kTile;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t r[32];
        if (n_it > 0) {
          tmem_ld_x32(tsrc + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int x = 0; x < 32; ++x) r[x] = 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 v;
          v.x = pack2<T>(__uint_as_float(r[8 * u + 0]) * mul, __uint_as_float(r[8 * u + 1]) * mul);
          v.y = pack2<T>(__uint_as_float(r[8 * u + 2]) * mul, __uint_as_float(r[8 * u + 3]) * mul);
          v.z = pack2<T>(__uint_as_float(r[8 * u + 4]) * mul, __uint_as_float(r[8 * u + 5]) * mul);
          v.w = pack2<T>(__uint_as_float(r[8 * u + 6]) * mul, __uint_as_float(r[8 * u + 7]) * mul);
          const int chunk = c * 4 + u;
          const int bx = chunk >> 3, cin = chunk & 7;
          *reinterpret_cast<uint4*>(stage + bx * S::kBox + row * 128 + ((cin ^ (row & 7)) << 4)) = v;
        }
      }
      if (warp == 4) BWD_CTA(8);   // TMEM drained, staged
      const int rows_left = KL - k0;
      if (packed && rows_left < 128) {
        // partial last block of a packed sequence: copy only its own rows (coalesced 16-byte stores)
        named_bar_sync(1 + half, 128);
        constexpr int kCPR = D / 8;
        T* obase = static_cast<T*>(half ? p.dk_ptr : p.dv_ptr) +
                   (static_cast<int64_t>(hk) * p.total_k + k_off + k0) * D;
        const int tid = wq * 32 + lane;
        for (int idx = tid; idx < rows_left * kCPR; idx += 128) {
          const int r = idx / kCPR, chunk = idx % kCPR;
          const int bx = chunk >> 3, cin = chunk & 7;
          const uint4 v = *reinterpret_cast<const uint4*>(stage + bx * S::kBox + r * 128 +
                                                          ((cin ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(r) * D + chunk * 8) = v;
        }
      } else {
        fence_proxy_async_smem();
        named_bar_sync(1 + half, 128);
        if (wq == 0 && lane == 0) {
#pragma unroll
          for (int bx = 0; bx < S::kNBox; ++bx)
            tma_store_3d(half ? &tm_dk : &tm_dv, stage + bx * S::kBox, bx * 64, k_off + k0, bh_kv);
          bulk_commit();
          if (warp == 4) BWD_CTA(9);
          bulk_wait_read<0>();
          if (warp == 4) BWD_CTA(10);
        }
      }
    }
  } else {
    // ================================ dQ drain warpgroup ===============================
    setmaxnreg_inc<152>();
    const int wq = warp & 3;
    const int row = wq * 32 + lane;  // query row within the block
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const bool issuer = (warp == 12 && lane == 0);
    int nred = 0;
    for (int it = 0; it < n_it; ++it) {
      const int bh_q = b * p.QH + hk * g + it / nqi;
      const int q0 = (i0 + it % nqi) * 128;
      mbar_wait(dq_full, it & 1);
      tc_fence_after();
      if (wq == 0) BWD_STAMP(it, 12);
      // pull the whole dQ_i tile into registers first so its TMEM columns (shared with dP^T)
      // are released as early as possible -- this read sits on the dS -> dQ -> dP^T(i+1) chain
      uint32_t r[D / 32][32];
#pragma unroll
      for (int c = 0; c < D / 32; ++c) tmem_ld_x32(tmem_base + lane_off + kColDP + c * 32, r[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);
      if (wq == 0) BWD_STAMP(it, 13);
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint8_t* stage = sdQ + (nred & 1) * 16384;
        if (issuer) bulk_wait_read<1>();  // the reduce that last read this buffer has finished
        named_bar_sync(3, 128);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint4 v = make_uint4(r[c][4 * u], r[c][4 * u + 1], r[c][4 * u + 2], r[c][4 * u + 3]);
          *reinterpret_cast<uint4*>(stage + row * 128 + ((u ^ (row & 7)) << 4)) = v;
        }
        fence_proxy_async_smem();
        named_bar_sync(3, 128);
        if (issuer) {
          tma_reduce_add_3d(&tm_dqa, stage, c * 32, q_off + q0, bh_q);
          bulk_commit();
        }
        ++nred;
      }
    }
    if (issuer) BWD_CTA(11);
    if (issuer) bulk_wait<0>();
    if (issuer) BWD_CTA(12);
  }

  // ---- teardown -------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
#ifdef NNOP_BWD_TRACE
  if (threadIdx.x == 0 && cta_slot >= 0) {
    g_bwd_cta_log[cta_slot * 16 + 5] = clock64();
    g_bwd_cta_log[cta_slot * 16 + 6] = blockIdx.x;
    g_bwd_cta_log[cta_slot * 16 + 7] = n_it;
  }
#endif
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// =========================================================================================
// Persistent variant (dense layout, no key padding mask): one CTA per SM walks a dynamic queue of
// (kv block, kv head, batch) tiles.  The plain kernel above spends ~9 % of a CTA's life outside
// its steady state (prologue, dK/dV epilogue, launch gap; scripts/trace_bwd_ctas.py) and with one
// CTA per SM nothing overlaps it.  Here the barrier phases, the Q_i / dO_i ring and TMEM run on
// across tiles:
//   * warp 3 is the scheduler: atomicAdd on a global tile counter (zeroed by the prep kernel), tile
//     ids broadcast through a two-slot shared-memory ring (tile_full / tile_empty);
//   * the TMA producer loads the next tile's V as soon as the last dP^T has read V (one step before
//     the tile ends), its first Q_i / dO_i as their ring slots free up, and K after the last dQ;
//   * the dK / dV epilogue moves to the dQ drain warpgroup (staging in its own 32 KB buffer), so the
//     compute warpgroups go straight on to the next tile's logits; the MMA warp waits on dv_empty /
//     dk_empty only before the next tile's first accumulating MMA overwrites the accumulators.
// Per tile the summation order is that of the plain kernel, so dK / dV agree bit for bit.
// =========================================================================================
struct PersistBars {
  enum : int {
    kTileFull = 0,   // [2] scheduler -> all roles
    kTileEmpty = 2,  // [2] all roles -> scheduler (14 arrivals)
    kKFull = 4, kVFull = 5, kKEmpty = 6, kVEmpty = 7,
    kQFull = 8,      // [2]
    kQEmpty = 10,    // [2]
    kDoFull = 12, kDoEmpty = 13,
    kSFull = 14, kPFull = 15, kDpFull = 16, kDsFull = 17, kDqFull = 18, kDqEmpty = 19,
    kDkDvFull = 20, kDvEmpty = 21, kDkEmpty = 22,
    kSchedGo = 23,   // producer -> scheduler: the current tile is nearly loaded, claim the next one
    kXFull = 24,     // DUO: the peer CTA's half of dQ_i has landed in this CTA's staging buffer (4 warp arrivals)
    kXFree = 25,     // DUO: the peer CTA's staging buffer may be overwritten (1 arrival, from the peer)
    kCount = 26
  };
};
static_assert(PersistBars::kCount + 4 <= BwdSmem<128>::kNumBars, "barrier area too small for the tile ring");

// DUO = true (launched as clusters of two CTAs): the pair shares one tile record (batch, kv head, PAIR of
// kv blocks 2*jp, 2*jp + 1); CTA r of the pair owns kv block 2*jp + r and both walk the same q blocks
// i >= 2*jp (CTA 1's first causal step lies above its diagonal: fully masked, contributes zeros).  Each
// CTA's dQ_i partial is halved by columns: the half the peer reduces goes straight from registers into
// the peer's staging buffer over distributed shared memory, the other half is summed with what the peer
// sent and leaves in ONE bulk reduce-add of half the width -- the fp32 read-modify-write traffic into L2
// (148 SMs x 64 KB per step, the 650-clk item of profiles/r01c_bwd_knockouts.txt) is cut in two.  The MMA
// chain never waits on the peer: the drain warpgroup releases dQ's TMEM columns before it exchanges.
template <typename T, int D, bool DUO = false>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_sm100_persist_kernel(const __grid_constant__ CUtensorMap tm_q,
                              const __grid_constant__ CUtensorMap tm_k,
                              const __grid_constant__ CUtensorMap tm_v,
                              const __grid_constant__ CUtensorMap tm_do,
                              const __grid_constant__ CUtensorMap tm_dk,
                              const __grid_constant__ CUtensorMap tm_dv,
                              const __grid_constant__ CUtensorMap tm_dqa, const BwdParams p,
                              int* __restrict__ tile_counter, const int n_tiles, const int nkv) {
  using S = BwdSmem<D>;
  using PB = PersistBars;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + S::kK;
  uint8_t* sV = smem + S::kV;
  uint8_t* sQ = smem + S::kQ;
  uint8_t* sdO = smem + S::kdO;
  uint8_t* sdS = smem + S::kdS;
  uint8_t* sdQ = smem + S::kdQs;
  float* s_lse = reinterpret_cast<float*>(smem + S::kStat);
  float* s_del = reinterpret_cast<float*>(smem + S::kStat + 1024);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + S::kNumBars);
  // tile ring: {batch element or packed sequence, kv block, kv head, 1 = tile / -1 = queue empty}
  volatile int* s_tile = reinterpret_cast<volatile int*>(bars + PersistBars::kCount);  // [2][4]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = p.QH / p.KH;
  const bool packed = !DUO && p.cu_q != nullptr;
  const uint32_t crank = DUO ? cluster_ctarank() : 0u;   // CTA of the pair (DUO), else 0

  // everything a role needs to know about a tile, derived from its ring record
  struct Tile {
    int b, j, hk, k0, QL, KL, q_off, k_off, st_off, i0, nqi, n_it, bh_kv;
  };
  auto make_tile = [&](int zb, int jrec, int hk) -> Tile {
    Tile t;
    const int j = DUO ? 2 * jrec + static_cast<int>(crank) : jrec;   // DUO: the record names the pair
    t.j = j; t.hk = hk; t.k0 = j * 128;
    if (packed) {
      t.q_off = p.cu_q[zb];
      t.QL = p.cu_q[zb + 1] - t.q_off;
      t.k_off = p.cu_k[zb];
      t.KL = p.cu_k[zb + 1] - t.k_off;
      t.st_off = packed_stat_row(t.q_off, zb);
      t.b = 0;
    } else {
      t.b = zb; t.QL = p.QL; t.KL = p.KL; t.q_off = 0; t.k_off = 0; t.st_off = 0;
    }
    const int nq = (t.QL + 127) >> 7;
    t.i0 = p.causal ? (DUO ? 2 * jrec : j) : 0;   // DUO: both CTAs walk the q blocks of the pair's first kv block
    t.nqi = nq > t.i0 ? nq - t.i0 : 0;
    t.n_it = t.nqi * g;
    t.bh_kv = t.b * p.KH + hk;
    return t;
  };
  // consumers of a tile record release its ring slot to the scheduler, which lives in CTA 0 of a pair
  auto release_tile_slot = [&](int slot) {
    if (DUO && crank != 0) mbar_arrive_cluster(bars + PB::kTileEmpty + slot, 0);
    else mbar_arrive(bars + PB::kTileEmpty + slot);
  };
  auto wait_tile_full = [&](int slot, uint32_t parity) {
    if constexpr (DUO) mbar_wait_cluster(bars + PB::kTileFull + slot, parity);   // record written by CTA 0 over DSMEM
    else mbar_wait(bars + PB::kTileFull + slot, parity);
  };

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("nnop: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dqa);
    tma_prefetch_desc(&tm_dk);
    tma_prefetch_desc(&tm_dv);
    for (int i = 0; i < PB::kCount; ++i) {
      uint32_t cnt = 1;
      if (i == PB::kTileEmpty || i == PB::kTileEmpty + 1) cnt = DUO ? 28 : 14;  // TMA + MMA + 8 compute + 4 drain (per CTA)
      if (i == PB::kPFull || i == PB::kDsFull) cnt = 8;              // one arrival per compute warp
      if (i == PB::kDqEmpty || i == PB::kDvEmpty || i == PB::kDkEmpty || i == PB::kXFull) cnt = 4;  // per drain warp
      mbar_init(bars + i, cnt);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (DUO) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 256 + D;

  // every consumer role fetches the n-th tile record the same way (false = queue empty)
  auto next_tile = [&](int n, Tile& t) -> bool {
    const int slot = n & 1;
    wait_tile_full(slot, (n >> 1) & 1);
    const int zb = s_tile[4 * slot], j = s_tile[4 * slot + 1], hk = s_tile[4 * slot + 2];
    const int flag = s_tile[4 * slot + 3];
    __syncwarp();
    if (lane == 0) release_tile_slot(slot);
    if (flag < 0) return false;
    t = make_tile(zb, j, hk);
    return true;
  };

  if (warp < 4) {
    if constexpr (DUO) setmaxnreg_dec<64>(); else setmaxnreg_dec<88>();   // DUO: the drain warpgroup holds all of dQ_i while it exchanges
    if (warp == 3 && crank == 0) {
      // ================================ tile scheduler ===============================
      // The whole warp runs it (uniformly); lane 0 claims and publishes.  Dense: tile t = ((b * KH +
      // hk) * nkv + j).  Packed: sequence z by binary search in tile_pre, then hk-major, j fastest.
      const int total = packed ? p.tile_pre[p.nseq] : n_tiles;
      for (int n = 0;; ++n) {
        const int slot = n & 1;
        // a tile is claimed only when this CTA is about to need it: tiles claimed early would sit
        // in the ring while other CTAs run dry at the end of the queue
        if (n > 0) mbar_wait(bars + PB::kSchedGo, (n - 1) & 1);
        mbar_wait(bars + PB::kTileEmpty + slot, ((n >> 1) & 1) ^ 1);
        int zb = 0, j = 0, hk = 0, flag = -1;
        for (;;) {
          int t = 0;
          if (lane == 0) t = atomicAdd(tile_counter, 1);
          t = __shfl_sync(0xffffffffu, t, 0);
          if (t >= total) break;
          if (packed) {
            int lo = 0, hi = p.nseq;  // largest z with tile_pre[z] <= t
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if (p.tile_pre[mid] <= t) lo = mid; else hi = mid;
            }
            zb = lo;
            const int tl = t - p.tile_pre[zb];
            const int nkv_z = (p.tile_pre[zb + 1] - p.tile_pre[zb]) / p.KH;
            hk = tl / nkv_z;
            j = tl - hk * nkv_z;
          } else if (p.lpt_group > 1) {
            // unit by unit, the last unit's heaviest tile is claimed when the queue is ~99 % drained and
            // then runs alone; group-wise heaviest-first leaves only light tiles for the end
            const int nbh = n_tiles / nkv;
            const int per = p.lpt_group * nkv;
            const int g0 = (t / per) * p.lpt_group;
            const int gsz = min(p.lpt_group, nbh - g0);
            const int idx = t - (t / per) * per;
            j = idx / gsz;
            const int u = g0 + idx - j * gsz;
            hk = u % p.KH;
            zb = u / p.KH;
          } else {
            j = t % nkv;
            const int u = t / nkv;
            hk = u % p.KH;
            zb = u / p.KH;
          }
          const Tile ti = make_tile(zb, j, hk);
          if (ti.n_it > 0) { flag = 1; break; }
          // no visible query for this kv block (packed sequence without queries): dK = dV = 0, written
          // here with plain stores; the tile is not published
          {
            const int rows = min(128, ti.KL - ti.k0);
            constexpr int kCPR = D / 8;
            const int64_t base = (static_cast<int64_t>(hk) * p.total_k + ti.k_off + ti.k0) * D;
            for (int idx = lane; idx < rows * kCPR; idx += 32) {
              reinterpret_cast<uint4*>(static_cast<T*>(p.dk_ptr) + base)[idx] = make_uint4(0u, 0u, 0u, 0u);
              reinterpret_cast<uint4*>(static_cast<T*>(p.dv_ptr) + base)[idx] = make_uint4(0u, 0u, 0u, 0u);
            }
          }
        }
        if (lane == 0) {
          s_tile[4 * slot] = zb; s_tile[4 * slot + 1] = j; s_tile[4 * slot + 2] = hk; s_tile[4 * slot + 3] = flag;
          if constexpr (DUO) {   // the same record into the peer's ring, then its tile_full (release at cluster scope)
            st_cluster_v4(mapa_u32(smem_u32(const_cast<int*>(s_tile) + 4 * slot), 1), static_cast<uint32_t>(zb),
                          static_cast<uint32_t>(j), static_cast<uint32_t>(hk), static_cast<uint32_t>(flag));
            mbar_arrive_cluster(bars + PB::kTileFull + slot, 1);
          }
          mbar_arrive(bars + PB::kTileFull + slot);
        }
        if (flag < 0) break;
      }
    } else if (warp == 0) {
      // ================================ TMA producer =================================
      if (lane == 0) {
        int gs = 0;
        for (int tl = 0;; ++tl) {
          const int slot = tl & 1;
          wait_tile_full(slot, (tl >> 1) & 1);
          const int zb = s_tile[4 * slot], tj = s_tile[4 * slot + 1], thk = s_tile[4 * slot + 2];
          const int flag = s_tile[4 * slot + 3];
          release_tile_slot(slot);
          if (flag < 0) break;
          const Tile ti = make_tile(zb, tj, thk);
          const int hk = ti.hk, b = ti.b, k0 = ti.k0, bh_kv = ti.bh_kv, i0 = ti.i0, nqi = ti.nqi, n_it = ti.n_it;
          const int q_off = ti.q_off, k_off = ti.k_off, st_off = ti.st_off;
          auto load_q = [&](int it) {
            const int gi = gs + it;
            const int s = gi & 1;
            const int bh_q = b * p.QH + hk * g + it / nqi;
            const int q0 = (i0 + it % nqi) * 128;
            mbar_wait(bars + PB::kQEmpty + s, ((gi >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx(bars + PB::kQFull + s, S::kTile + 1024);
#pragma unroll
            for (int bx = 0; bx < S::kNBox; ++bx)
              tma_load_3d(sQ + s * S::kTile + bx * S::kBox, &tm_q, bars + PB::kQFull + s, bx * 64, q_off + q0, bh_q);
            const int64_t soff = static_cast<int64_t>(bh_q) * p.QLp + st_off + q0;
            bulk_load_1d(s_lse + s * 128, p.lse2p + soff, 512, bars + PB::kQFull + s);
            bulk_load_1d(s_del + s * 128, p.deltap + soff, 512, bars + PB::kQFull + s);
          };
          auto load_do = [&](int it) {
            const int gi = gs + it;
            const int bh_q = b * p.QH + hk * g + it / nqi;
            const int q0 = (i0 + it % nqi) * 128;
            mbar_wait(bars + PB::kDoEmpty, (gi & 1) ^ 1);
            mbar_arrive_expect_tx(bars + PB::kDoFull, S::kTile);
#pragma unroll
            for (int bx = 0; bx < S::kNBox; ++bx)
              tma_load_3d(sdO + bx * S::kBox, &tm_do, bars + PB::kDoFull, bx * 64, q_off + q0, bh_q);
          };
          const int go_at = n_it > 2 ? n_it - 2 : 0;  // step whose loads trigger the next claim
          if (go_at == 0 && crank == 0) mbar_arrive(bars + PB::kSchedGo);
          // in the order the previous tile releases them: V, Q ring slot, dO, K
          mbar_wait(bars + PB::kVEmpty, (tl & 1) ^ 1);
          mbar_arrive_expect_tx(bars + PB::kVFull, S::kTile);
#pragma unroll
          for (int bx = 0; bx < S::kNBox; ++bx)
            tma_load_3d(sV + bx * S::kBox, &tm_v, bars + PB::kVFull, bx * 64, k_off + k0, bh_kv);
          load_q(0);
          load_do(0);
          mbar_wait(bars + PB::kKEmpty, (tl & 1) ^ 1);
          mbar_arrive_expect_tx(bars + PB::kKFull, S::kTile);
#pragma unroll
          for (int bx = 0; bx < S::kNBox; ++bx)
            tma_load_3d(sK + bx * S::kBox, &tm_k, bars + PB::kKFull, bx * 64, k_off + k0, bh_kv);
          if (n_it > 1) load_q(1);
          for (int it = 1; it < n_it; ++it) {
            if (it == go_at && crank == 0) mbar_arrive(bars + PB::kSchedGo);
            load_do(it);
            if (it + 1 < n_it) load_q(it + 1);
          }
          gs += n_it;
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ===================================
      constexpr bool BF = is_bf16<T>::value;
      constexpr uint32_t id_kk = make_idesc_f16(128, 128, BF, false, false);  // S^T, dP^T
      constexpr uint32_t id_tv = make_idesc_f16(128, D, BF, false, true);     // dV (A in TMEM), dK
      constexpr uint32_t id_mm = make_idesc_f16(128, D, BF, true, true);      // dQ
      const uint32_t tm = uniform_u32(tmem_base);
      const uint32_t sbase = uniform_u32(smem_u32(smem));
      const uint64_t kmaj = make_smem_desc_sw128(sbase, 16, 1024);
      const uint64_t mnmaj = make_smem_desc_sw128(sbase, S::kBox, 1024);
      const uint32_t k_lo = desc_lo(kmaj), k_hi = desc_hi(kmaj);
      const uint32_t m_lo = desc_lo(mnmaj), m_hi = desc_hi(mnmaj);
      auto mma_kk = [&](uint32_t dcol, uint32_t a0, uint32_t b0) {
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t off = (ks >> 2) * S::kBox + (ks & 3) * 32;
            umma_ss_lo(tm + dcol, k_lo, (a0 + off) >> 4, k_hi, k_lo, (b0 + off) >> 4, k_hi, id_kk,
                       ks > 0 ? 1u : 0u);
          }
        }
      };
      auto commit = [&](int bar) {
        if (elect_one()) tc_commit(bars + bar);
      };
      int gs = 0;
      for (int tl = 0;; ++tl) {
        Tile ti;
        if (!next_tile(tl, ti)) break;
        [[maybe_unused]] const int j = ti.j;  // (trace builds log it)
        const int n_it = ti.n_it;
        PT_STAMP(0);
        // dP^T(0) = V dO_0^T: its TMEM columns hold the previous tile's last dQ until drained
        mbar_wait(bars + PB::kVFull, tl & 1);
        mbar_wait(bars + PB::kDoFull, gs & 1);
        if (gs > 0) mbar_wait(bars + PB::kDqEmpty, (gs - 1) & 1);
        tc_fence_after();
        PT_STAMP(1);
        mma_kk(kColDP, S::kV, S::kdO);
        commit(PB::kDpFull);
        if (n_it == 1) commit(PB::kVEmpty);
        // S^T(0) = K Q_0^T
        mbar_wait(bars + PB::kKFull, tl & 1);
        mbar_wait(bars + PB::kQFull + (gs & 1), (gs >> 1) & 1);
        tc_fence_after();
        PT_STAMP(2);
        mma_kk(kColS, S::kK, S::kQ + static_cast<uint32_t>((gs & 1) * S::kTile));
        commit(PB::kSFull);
        for (int it = 0; it < n_it; ++it) {
          const int gi = gs + it;
          const int s = gi & 1;
          const uint32_t acc = it > 0 ? 1u : 0u;
          const uint32_t qoff = static_cast<uint32_t>(s * S::kTile);
          // dV += P^T dO_i
          if (it == 0) mbar_wait(bars + PB::kDvEmpty, (tl & 1) ^ 1);  // previous tile's dV read out
          mbar_wait(bars + PB::kPFull, gi & 1);
          tc_fence_after();
          if (it == 0) PT_STAMP(3);
          if (it == 1) PT_STAMP(9);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_ts_lo(tm + kColDV, tm + kColS + (ks >> 2) * 64 + (ks & 3) * 8, m_lo,
                         (S::kdO + ks * 2048) >> 4, m_hi, id_tv, (acc | (ks > 0)) ? 1u : 0u);
          }
          commit(PB::kDoEmpty);
          // S^T(i+1)
          if (it + 1 < n_it) {
            mbar_wait(bars + PB::kQFull + (s ^ 1), ((gi + 1) >> 1) & 1);
            tc_fence_after();
            mma_kk(kColS, S::kK, S::kQ + static_cast<uint32_t>((s ^ 1) * S::kTile));
            commit(PB::kSFull);
          }
#ifndef NNOP_BWD_DQ_FIRST
          // dK(i) before dQ(i): the Q stage is released one MMA earlier, and the load of Q_{i+2} into it
          // sits on the S^T(i+2) critical path (measured: steady step 3 952 -> 3 870 clk); dQ(i), and
          // with it dP^T(i+1), move back by one MMA, which the dS phase has the slack for
          mbar_wait(bars + PB::kDsFull, gi & 1);
          tc_fence_after();
          if (it == 0) {
            mbar_wait(bars + PB::kDkEmpty, (tl & 1) ^ 1);
            tc_fence_after();
            PT_STAMP(5);
          }
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t off = (ks >> 2) * S::kBox + (ks & 3) * 32;
              umma_ss_lo(tm + kColDK, k_lo, (S::kdS + off) >> 4, k_hi, m_lo, (S::kQ + qoff + ks * 2048) >> 4,
                         m_hi, id_tv, (acc | (ks > 0)) ? 1u : 0u);
            }
          }
          commit(PB::kQEmpty + s);
          if (it == 0) PT_STAMP(4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_ss_lo(tm + kColDP, m_lo, (S::kdS + ks * 2048) >> 4, m_hi, m_lo, (S::kK + ks * 2048) >> 4,
                         m_hi, id_mm, ks > 0 ? 1u : 0u);
          }
          commit(PB::kDqFull);
          if (it + 1 == n_it) commit(PB::kKEmpty);
#else
          // dQ_i = dS K_j
          mbar_wait(bars + PB::kDsFull, gi & 1);
          tc_fence_after();
          if (it == 0) PT_STAMP(4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma_ss_lo(tm + kColDP, m_lo, (S::kdS + ks * 2048) >> 4, m_hi, m_lo, (S::kK + ks * 2048) >> 4,
                         m_hi, id_mm, ks > 0 ? 1u : 0u);
          }
          commit(PB::kDqFull);
          if (it + 1 == n_it) commit(PB::kKEmpty);  // K_j is not read again
          // dK += dS^T Q_i
          if (it == 0) {
            mbar_wait(bars + PB::kDkEmpty, (tl & 1) ^ 1);  // previous tile's dK read out
            tc_fence_after();
            PT_STAMP(5);
          }
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t off = (ks >> 2) * S::kBox + (ks & 3) * 32;
              umma_ss_lo(tm + kColDK, k_lo, (S::kdS + off) >> 4, k_hi, m_lo, (S::kQ + qoff + ks * 2048) >> 4,
                         m_hi, id_tv, (acc | (ks > 0)) ? 1u : 0u);
            }
          }
          commit(PB::kQEmpty + s);
#endif
          // dP^T(i+1)
          if (it + 1 < n_it) {
            mbar_wait(bars + PB::kDoFull, (gi + 1) & 1);
            mbar_wait(bars + PB::kDqEmpty, gi & 1);
            tc_fence_after();
            mma_kk(kColDP, S::kV, S::kdO);
            commit(PB::kDpFull);
            if (it + 2 == n_it) commit(PB::kVEmpty);  // V_j is not read again
          }
        }
        commit(PB::kDkDvFull);
        PT_STAMP(6);
#ifdef NNOP_BWD_TRACE
        if (blockIdx.x == 0 && lane == 0 && tl < 1024) {
          g_bwd_cta_log[tl * 16 + 7] = n_it;
          g_bwd_cta_log[tl * 16 + 8] = j;
          g_bwd_cta_n = tl + 1;
        }
#endif
        gs += n_it;
      }
    }
  } else if (warp < 12) {
    // ================================ compute warpgroups ===============================
    setmaxnreg_inc<136>();
    const int half = (warp - 4) >> 2;  // which 64 q columns
    const int wq = warp & 3;
    const int row = wq * 32 + lane;    // key row within the block
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const int c0 = half * 64;
    const float sl2 = p.scale_log2;
    int gs = 0;
    for (int tl = 0;; ++tl) {
      Tile ti;
      if (!next_tile(tl, ti)) break;
      const int j = ti.j, i0 = ti.i0, nqi = ti.nqi, n_it = ti.n_it;
      const bool key_dead = packed && ti.k0 + row >= ti.KL;  // this row belongs to the next packed sequence
      for (int it = 0; it < n_it; ++it) {
        const int gi = gs + it;
        const int s = gi & 1;
        const int i = i0 + it % nqi;
        // ---- P^T ----
        mbar_wait(bars + PB::kQFull + s, (gi >> 1) & 1);  // lse2 / delta of this stage have landed
        mbar_wait(bars + PB::kSFull, gi & 1);
        tc_fence_after();
        if (it == 0 && warp == 4) PT_STAMP(13);
        uint32_t sr[2][32];
        tmem_ld_x32(tmem_base + lane_off + kColS + c0, sr[0]);
        tmem_ld_x32(tmem_base + lane_off + kColS + c0 + 32, sr[1]);
        tmem_ld_wait();
        float pf[64];
        const float4* l4 = reinterpret_cast<const float4*>(s_lse + s * 128 + c0);
#pragma unroll
        for (int u4 = 0; u4 < 16; ++u4) {
#ifdef NNOP_BWD_NO_STATS   // timing experiment only: what do the broadcast LDS of lse2 / delta cost?
          const float4 l = make_float4(sl2, sl2, sl2, sl2);
#else
          const float4 l = l4[u4];
#endif
          pf[4 * u4 + 0] = fast_exp2(fmaf(__uint_as_float(sr[u4 >> 3][(4 * u4 + 0) & 31]), sl2, -l.x));
          pf[4 * u4 + 1] = fast_exp2(fmaf(__uint_as_float(sr[u4 >> 3][(4 * u4 + 1) & 31]), sl2, -l.y));
          pf[4 * u4 + 2] = fast_exp2(fmaf(__uint_as_float(sr[u4 >> 3][(4 * u4 + 2) & 31]), sl2, -l.z));
          pf[4 * u4 + 3] = fast_exp2(fmaf(__uint_as_float(sr[u4 >> 3][(4 * u4 + 3) & 31]), sl2, -l.w));
        }
        if (p.causal && i == j) {  // diagonal block: key k0+row is visible to query q0+c iff row <= c
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (row > c0 + c) pf[c] = 0.f;
        }
        if (DUO && p.causal && i < j) {  // CTA 1 of a pair, first step: the whole block lies above the diagonal
#pragma unroll
          for (int c = 0; c < 64; ++c) pf[c] = 0.f;
        }
        if (key_dead) {
#pragma unroll
          for (int c = 0; c < 64; ++c) pf[c] = 0.f;
        }
        {
          uint32_t pk[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) pk[c] = pack2<T>(pf[2 * c], pf[2 * c + 1]);
          tmem_st_x32(tmem_base + lane_off + kColS + half * 64, pk);  // inside this half's own S^T columns
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + PB::kPFull);
        if (it == 0 && warp == 4) PT_STAMP(12);
        // ---- dS^T ----
        mbar_wait(bars + PB::kDpFull, gi & 1);
        tc_fence_after();
        tmem_ld_x32(tmem_base + lane_off + kColDP + c0, sr[0]);
        tmem_ld_x32(tmem_base + lane_off + kColDP + c0 + 32, sr[1]);
        tmem_ld_wait();
        const float4* d4 = reinterpret_cast<const float4*>(s_del + s * 128 + c0);
        uint8_t* drow = sdS + half * S::kBox + row * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {  // 8 columns = one 16-byte chunk
#ifdef NNOP_BWD_NO_STATS
          const float4 da = make_float4(sl2, sl2, sl2, sl2), db = da;
#else
          const float4 da = d4[2 * ch], db = d4[2 * ch + 1];
#endif
          const float dl[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
          float ds[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = 8 * ch + e;
            ds[e] = pf[c] * (__uint_as_float(sr[c >> 5][c & 31]) - dl[e]);
          }
          uint4 v;
          v.x = pack2<T>(ds[0], ds[1]);
          v.y = pack2<T>(ds[2], ds[3]);
          v.z = pack2<T>(ds[4], ds[5]);
          v.w = pack2<T>(ds[6], ds[7]);
#ifndef NNOP_BWD_NO_DS   // (timing experiments: knock out the dS^T shared-memory stores)
          *reinterpret_cast<uint4*>(drow + ((ch ^ (row & 7)) << 4)) = v;
#else
          if (v.x == 0x12345678u) *reinterpret_cast<uint4*>(drow) = v;
#endif
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + PB::kDsFull);
      }
      gs += n_it;
    }
  } else {
    // ========================= dQ drain + dK / dV epilogue warpgroup ====================
    if constexpr (DUO) setmaxnreg_inc<176>(); else setmaxnreg_inc<152>();
    const int wq = warp & 3;
    const int row = wq * 32 + lane;  // query row (dQ) / key row (dK, dV) within the block
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const bool issuer = (warp == 12 && lane == 0);
    int nred = 0;
    int gs = 0;
    for (int tl = 0;; ++tl) {
      Tile ti;
      if (!next_tile(tl, ti)) break;
      const int hk = ti.hk, b = ti.b, k0 = ti.k0, bh_kv = ti.bh_kv, i0 = ti.i0, nqi = ti.nqi, n_it = ti.n_it;
      const int q_off = ti.q_off, k_off = ti.k_off;
      for (int it = 0; it < n_it; ++it) {
        const int gi = gs + it;
        const int bh_q = b * p.QH + hk * g + it / nqi;
        const int q0 = (i0 + it % nqi) * 128;
        mbar_wait(bars + PB::kDqFull, gi & 1);
        tc_fence_after();
        uint32_t r[D / 32][32];
#pragma unroll
        for (int c = 0; c < D / 32; ++c) tmem_ld_x32(tmem_base + lane_off + kColDP + c * 32, r[c]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + PB::kDqEmpty);
#ifndef NNOP_BWD_NO_DQ   // (timing experiments: knock out the dQ staging + bulk reduce)
        if constexpr (DUO) {
          // This CTA reduces columns [crank * D/2, +D/2) of dQ_i for BOTH kv blocks of the pair; the other
          // half of its partial goes to the peer.  X = the 32 KB staging buffer: landing area for the peer's
          // half, then (summed in place) the source of the bulk reduce-add.
          constexpr int NCH = D / 64;   // 32-column (16 KB) boxes per half
          const uint32_t peer = crank ^ 1u;
          if (issuer) {                 // my X has been read by its last bulk op: the peer may overwrite it
            bulk_wait_read<0>();
            mbar_arrive_cluster(bars + PB::kXFree, peer);
          }
          mbar_wait_cluster(bars + PB::kXFree, gi & 1);   // ... and the peer's X is free for my half
          const uint32_t xpeer = mapa_u32(smem_u32(sdQ), peer);
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int u4 = 0; u4 < 8; ++u4) {
              const uint32_t off = c * 16384 + row * 128 + ((u4 ^ (row & 7)) << 4);
              if (crank == 0)   // (static register indices in both branches: no local-memory indexing)
                st_cluster_v4(xpeer + off, r[NCH + c][4 * u4], r[NCH + c][4 * u4 + 1], r[NCH + c][4 * u4 + 2],
                              r[NCH + c][4 * u4 + 3]);
              else
                st_cluster_v4(xpeer + off, r[c][4 * u4], r[c][4 * u4 + 1], r[c][4 * u4 + 2], r[c][4 * u4 + 3]);
            }
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(bars + PB::kXFull, peer);
          mbar_wait_cluster(bars + PB::kXFull, gi & 1);   // the peer's half of MY columns has landed in X
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int u4 = 0; u4 < 8; ++u4) {
              float4* xp = reinterpret_cast<float4*>(sdQ + c * 16384 + row * 128 + ((u4 ^ (row & 7)) << 4));
              float4 v = *xp;
              if (crank == 0) {
                v.x += __uint_as_float(r[c][4 * u4]); v.y += __uint_as_float(r[c][4 * u4 + 1]);
                v.z += __uint_as_float(r[c][4 * u4 + 2]); v.w += __uint_as_float(r[c][4 * u4 + 3]);
              } else {
                v.x += __uint_as_float(r[NCH + c][4 * u4]); v.y += __uint_as_float(r[NCH + c][4 * u4 + 1]);
                v.z += __uint_as_float(r[NCH + c][4 * u4 + 2]); v.w += __uint_as_float(r[NCH + c][4 * u4 + 3]);
              }
              *xp = v;
            }
          fence_proxy_async_smem();
          named_bar_sync(3, 128);
          if (issuer) {
#pragma unroll
            for (int c = 0; c < NCH; ++c)
              tma_reduce_add_3d(&tm_dqa, sdQ + c * 16384, static_cast<int>(crank) * (D / 2) + c * 32, q_off + q0, bh_q);
            bulk_commit();
          }
        } else {
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
          uint8_t* stage = sdQ + (nred & 1) * 16384;
          if (issuer) {
            // the bulk op that last read this buffer has finished (after an epilogue: the dK store,
            // which reads both buffers)
            if (it == 0 && c == 0) bulk_wait_read<0>(); else bulk_wait_read<1>();
          }
          named_bar_sync(3, 128);
#pragma unroll
          for (int u4 = 0; u4 < 8; ++u4) {
            const uint4 v = make_uint4(r[c][4 * u4], r[c][4 * u4 + 1], r[c][4 * u4 + 2], r[c][4 * u4 + 3]);
#ifndef NNOP_BWD_NO_DQ_STS
            *reinterpret_cast<uint4*>(stage + row * 128 + ((u4 ^ (row & 7)) << 4)) = v;
#else
            if (v.x == 0x12345678u) *reinterpret_cast<uint4*>(stage) = v;
#endif
          }
          fence_proxy_async_smem();
          named_bar_sync(3, 128);
#ifndef NNOP_BWD_NO_DQ_RED
          if (issuer) {
            tma_reduce_add_3d(&tm_dqa, stage, c * 32, q_off + q0, bh_q);
            bulk_commit();
          }
#endif
          ++nred;
        }
        }
#endif
      }
      gs += n_it;
      // ---- epilogue: dV then dK -> 16-bit -> swizzled staging (the dQ buffers) -> TMA store ----
      // Each accumulator is pulled into registers in one go and released at once: the next tile's
      // first dV / dK MMAs wait on dv_empty / dk_empty, the staging and the store do not.
      if (warp == 12) PT_STAMP(14);
      mbar_wait(bars + PB::kDkDvFull, tl & 1);
      tc_fence_after();
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint32_t tsrc = tmem_base + lane_off + (which ? kColDK : kColDV);
        const float mul = which ? p.scale : 1.f;
        uint32_t r[D / 32][32];
#pragma unroll
        for (int c = 0; c < D / 32; ++c) tmem_ld_x32(tsrc + c * 32, r[c]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + (which ? PB::kDkEmpty : PB::kDvEmpty));
        if (warp == 12) PT_STAMP(10 + which);
        if (issuer) bulk_wait_read<0>();  // staging free: earlier dQ reduces / the dV store have been read
        named_bar_sync(3, 128);
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
#pragma unroll
          for (int u4 = 0; u4 < 4; ++u4) {
            uint4 v;
            v.x = pack2<T>(__uint_as_float(r[c][8 * u4 + 0]) * mul, __uint_as_float(r[c][8 * u4 + 1]) * mul);
            v.y = pack2<T>(__uint_as_float(r[c][8 * u4 + 2]) * mul, __uint_as_float(r[c][8 * u4 + 3]) * mul);
            v.z = pack2<T>(__uint_as_float(r[c][8 * u4 + 4]) * mul, __uint_as_float(r[c][8 * u4 + 5]) * mul);
            v.w = pack2<T>(__uint_as_float(r[c][8 * u4 + 6]) * mul, __uint_as_float(r[c][8 * u4 + 7]) * mul);
            const int chunk = c * 4 + u4;
            const int bx = chunk >> 3, cin = chunk & 7;
            *reinterpret_cast<uint4*>(sdQ + bx * S::kBox + row * 128 + ((cin ^ (row & 7)) << 4)) = v;
          }
        }
        const int rows_left = ti.KL - k0;
        if (packed && rows_left < 128) {
          // partial last block of a packed sequence: copy only its own rows (coalesced 16-byte stores);
          // the barrier that opens the next use of the staging buffer also closes these reads
          named_bar_sync(3, 128);
          constexpr int kCPR = D / 8;
          T* obase = static_cast<T*>(which ? p.dk_ptr : p.dv_ptr) +
                     (static_cast<int64_t>(hk) * p.total_k + k_off + k0) * D;
          for (int idx = row; idx < rows_left * kCPR; idx += 128) {
            const int r2 = idx / kCPR, chunk = idx % kCPR;
            const int bx = chunk >> 3, cin = chunk & 7;
            const uint4 v = *reinterpret_cast<const uint4*>(sdQ + bx * S::kBox + r2 * 128 + ((cin ^ (r2 & 7)) << 4));
            *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(r2) * D + chunk * 8) = v;
          }
        } else {
          fence_proxy_async_smem();
          named_bar_sync(3, 128);
          if (issuer) {
#pragma unroll
            for (int bx = 0; bx < S::kNBox; ++bx)
              tma_store_3d(which ? &tm_dk : &tm_dv, sdQ + bx * S::kBox, bx * 64, k_off + k0, bh_kv);
            bulk_commit();
          }
        }
      }
    }
    if (issuer) bulk_wait<0>();
  }

  // ---- teardown -------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if constexpr (DUO) cluster_sync_all();   // no CTA leaves while its peer may still write into it / arrive on it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// =========================================================================================
// CTA-pair variant (E = 128, dense layout): a cluster of two CTAs owns kv blocks (j, j+1) of one
// kv head and walks the SAME q blocks in lockstep.  S^T, dP^T, dV and dK become M=256
// cta_group::2 MMAs issued by the leader CTA: each CTA supplies its own 128 key rows of A and only
// HALF of the Q_i / dO_i operand (64 q rows for the K-major uses, 64 head-dim columns for the
// MN-major uses), which cuts the shared-memory port traffic of a step from 512 KB to 448 KB per CTA
// (that port, not the tensor pipe, bounds this kernel).  dQ_i = dS K_j stays a
// per-CTA cta_group::1 MMA (its two partial products have different B operands).
// Cross-CTA protocol: TMA loads of both CTAs count on the leader's barriers; the compute / drain
// warpgroups of CTA 1 arrive remotely on the leader's p_full / ds_pair / dq_empty; the leader's
// commits are multicast to the same-offset barriers of both CTAs.
// =========================================================================================
struct PairSmem {
  static constexpr int D = 128;
  static constexpr int kTile = 128 * D * 2;   // K_j, V_j (32 KB each)
  static constexpr int kBox = 128 * 64 * 2;   // 128 rows x 64 columns
  static constexpr int kHalf = 16384;         // half of a Q_i / dO_i tile
  static constexpr int kK = 0;
  static constexpr int kV = kK + kTile;
  static constexpr int kQA = kV + kTile;       // 2 stages x (64 q rows x 128 d): two 8 KB boxes each
  static constexpr int kQB = kQA + 2 * kHalf;  // 2 stages x (128 q rows x 64 d)
  static constexpr int kdOA = kQB + 2 * kHalf;
  static constexpr int kdOB = kdOA + kHalf;
  static constexpr int kdS = kdOB + kHalf;     // 128 x 128 16-bit, two boxes
  static constexpr int kdQs = kdS + 2 * kBox;  // 2 x (128 rows x 32 fp32)
  static constexpr int kStat = kdQs + 2 * 16384;
  static constexpr int kBar = kStat + 2048;
  static constexpr int kNumBars = 18;
  static constexpr int kTotal = kBar + kNumBars * 8 + 16;
};

template <typename T>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kBwdThreads, 1)
attn_bwd_pair_kernel(const __grid_constant__ CUtensorMap tm_q,     // box 64 x 128 rows
                     const __grid_constant__ CUtensorMap tm_q64,   // box 64 x 64 rows
                     const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v,
                     const __grid_constant__ CUtensorMap tm_do,
                     const __grid_constant__ CUtensorMap tm_do64,
                     const __grid_constant__ CUtensorMap tm_dk,
                     const __grid_constant__ CUtensorMap tm_dv,
                     const __grid_constant__ CUtensorMap tm_dqa, const BwdParams p) {
  using S = PairSmem;
  constexpr int D = 128;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem + S::kK;
  uint8_t* sV = smem + S::kV;
  uint8_t* sQA = smem + S::kQA;
  uint8_t* sQB = smem + S::kQB;
  uint8_t* sdOA = smem + S::kdOA;
  uint8_t* sdOB = smem + S::kdOB;
  uint8_t* sdS = smem + S::kdS;
  uint8_t* sdQ = smem + S::kdQs;
  float* s_lse = reinterpret_cast<float*>(smem + S::kStat);         // [2][128]
  float* s_del = reinterpret_cast<float*>(smem + S::kStat + 1024);  // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBar);
  uint64_t* kv_full = bars + 0;     // leader: K, V of both CTAs landed
  uint64_t* q_full = bars + 1;      // [2] leader: Q halves of both CTAs landed
  uint64_t* q_empty = bars + 3;     // [2] both (multicast commit)
  uint64_t* do_full = bars + 5;     // leader
  uint64_t* do_empty = bars + 6;    // both (multicast)
  uint64_t* s_full = bars + 7;      // both (multicast)
  uint64_t* p_full = bars + 8;      // leader: 16 warp arrivals (8 local + 8 remote)
  uint64_t* dp_full = bars + 9;     // both (multicast)
  uint64_t* ds_local = bars + 10;   // per CTA: 8 warp arrivals (gates the local dQ MMA)
  uint64_t* ds_pair = bars + 11;    // leader: 16 warp arrivals (gates the pair's dK MMA)
  uint64_t* dq_full = bars + 12;    // per CTA
  uint64_t* dq_empty = bars + 13;   // leader: 8 warp arrivals (4 local + 4 remote)
  uint64_t* dkdv_full = bars + 14;  // both (multicast)
  uint64_t* stat_full = bars + 15;  // [2] per CTA: lse2 / delta of the stage landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + S::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
#ifdef NNOP_BWD_TRACE
  const bool tr = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
#endif

  // ---- work assignment (identical trip counts in both CTAs) ------------------------------
  const int j = blockIdx.x;          // this CTA's kv block
  const int j0 = j & ~1;             // the pair's first block
  const int k0 = j * 128;
  const int hk = blockIdx.y, b = blockIdx.z;
  const int QL = p.QL, KL = p.KL;
  const int g = p.QH / p.KH;
  const int bh_kv = b * p.KH + hk;
  const int nq = (QL + 127) >> 7;
  const int i0 = p.causal ? j0 : 0;
  const int nqi = nq > i0 ? nq - i0 : 0;
  const int n_it = nqi * g;
  bool key_keep = true;
  if (p.kpad) {
    const int kr = k0 + (threadIdx.x & 127);
    key_keep = kr < KL && p.kpad[static_cast<int64_t>(b) * p.KL + kr] != 0;
  }

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("nnop: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_q64);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_do64);
    tma_prefetch_desc(&tm_dqa);
    mbar_init(kv_full, 1);
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    mbar_init(&q_empty[0], 1);
    mbar_init(&q_empty[1], 1);
    mbar_init(do_full, 1);
    mbar_init(do_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 16);
    mbar_init(dp_full, 1);
    mbar_init(ds_local, 8);
    mbar_init(ds_pair, 16);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 8);
    mbar_init(dkdv_full, 1);
    mbar_init(&stat_full[0], 1);
    mbar_init(&stat_full[1], 1);
    fence_mbar_init();
  }
  cluster_sync_all();  // barriers of both CTAs exist before any remote arrive / TMA completion
  if (warp == 2) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t kColS = 0, kColDP = 128, kColDV = 256, kColDK = 256 + D;

  if (warp < 4) {
    setmaxnreg_dec<88>();
    if (warp == 0 && lane == 0 && n_it > 0) {
      // ================================ TMA producer (both CTAs) =====================
      if (leader) mbar_arrive_expect_tx(kv_full, 4 * S::kTile);
#pragma unroll
      for (int bx = 0; bx < 2; ++bx) {
        tma_load_3d_pair(sK + bx * S::kBox, &tm_k, kv_full, bx * 64, k0, bh_kv);
        tma_load_3d_pair(sV + bx * S::kBox, &tm_v, kv_full, bx * 64, k0, bh_kv);
      }
      auto load_q = [&](int it) {
        const int s = it & 1;
        const int bh_q = b * p.QH + hk * g + it / nqi;
        const int q0 = (i0 + it % nqi) * 128;
        mbar_wait(&q_empty[s], ((it >> 1) & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(&q_full[s], 4 * S::kHalf);
        // K-major half: q rows [64*rank, 64*rank+64), all of E (two 8 KB boxes)
#pragma unroll
        for (int bx = 0; bx < 2; ++bx)
          tma_load_3d_pair(sQA + s * S::kHalf + bx * 8192, &tm_q64, &q_full[s], bx * 64,
                           q0 + 64 * static_cast<int>(rank), bh_q);
        // MN-major half: all 128 q rows, E columns [64*rank, 64*rank+64)
        tma_load_3d_pair(sQB + s * S::kHalf, &tm_q, &q_full[s], 64 * static_cast<int>(rank), q0, bh_q);
        const int64_t soff = static_cast<int64_t>(bh_q) * p.QLp + q0;
        mbar_arrive_expect_tx(&stat_full[s], 1024);
        bulk_load_1d(s_lse + s * 128, p.lse2p + soff, 512, &stat_full[s]);
        bulk_load_1d(s_del + s * 128, p.deltap + soff, 512, &stat_full[s]);
      };
      auto load_do = [&](int it) {
        const int bh_q = b * p.QH + hk * g + it / nqi;
        const int q0 = (i0 + it % nqi) * 128;
        mbar_wait(do_empty, (it & 1) ^ 1);
        if (leader) mbar_arrive_expect_tx(do_full, 4 * S::kHalf);
#pragma unroll
        for (int bx = 0; bx < 2; ++bx)
          tma_load_3d_pair(sdOA + bx * 8192, &tm_do64, do_full, bx * 64, q0 + 64 * static_cast<int>(rank), bh_q);
        tma_load_3d_pair(sdOB, &tm_do, do_full, 64 * static_cast<int>(rank), q0, bh_q);
      };
      load_q(0);
      load_do(0);
      if (n_it > 1) load_q(1);
      for (int it = 1; it < n_it; ++it) {
        load_do(it);
        if (it + 1 < n_it) load_q(it + 1);
      }
    } else if (warp == 1 && n_it > 0) {
      // ================================ MMA issuer ===================================
      constexpr bool BF = is_bf16<T>::value;
      constexpr uint32_t id2_kk = make_idesc_f16(256, 128, BF, false, false);  // S^T, dP^T (pair)
      constexpr uint32_t id2_tv = make_idesc_f16(256, D, BF, false, true);     // dV, dK (pair)
      constexpr uint32_t id_mm = make_idesc_f16(128, D, BF, true, true);       // dQ (per CTA)
      const uint32_t tm = uniform_u32(tmem_base);
      const uint32_t sbase = uniform_u32(smem_u32(smem));
      const uint64_t kmaj = make_smem_desc_sw128(sbase, 16, 1024);
      const uint64_t mnmaj = make_smem_desc_sw128(sbase, S::kBox, 1024);
      const uint32_t k_lo = desc_lo(kmaj), k_hi = desc_hi(kmaj);
      const uint32_t m_lo = desc_lo(mnmaj), m_hi = desc_hi(mnmaj);
      auto dq_local = [&]() {   // dQ_i = dS K_j on this CTA's tensor core
#ifdef NNOP_PAIR_NO_DQ
        if (false) {
#else
        if (elect_one()) {
#endif
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ss_lo(tm + kColDP, m_lo, (S::kdS + ks * 2048) >> 4, m_hi, m_lo, (S::kK + ks * 2048) >> 4,
                       m_hi, id_mm, ks > 0 ? 1u : 0u);
        }
        if (elect_one()) tc_commit(dq_full);
      };
      if (leader) {
        // D[256 x 128] = [A_cta0; A_cta1][256 x E] * B[128 x E]^T; B rows split 64 / 64 over the CTAs
        auto mma_kk = [&](uint32_t dcol, uint32_t a0, uint32_t b0) {
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
              const uint32_t aoff = (ks >> 2) * S::kBox + (ks & 3) * 32;
              const uint32_t boff = (ks >> 2) * 8192 + (ks & 3) * 32;
              umma2_ss_lo(tm + dcol, k_lo, (a0 + aoff) >> 4, k_hi, k_lo, (b0 + boff) >> 4, k_hi, id2_kk,
                          ks > 0 ? 1u : 0u);
            }
          }
        };
        auto commit = [&](uint64_t* bar) {
          if (elect_one()) tc_commit_pair(bar);
        };
        mbar_wait_cluster(kv_full, 0);
        mbar_wait_cluster(&q_full[0], 0);
        tc_fence_after();
        mma_kk(kColS, S::kK, S::kQA);
        commit(s_full);
        mbar_wait_cluster(do_full, 0);
        tc_fence_after();
        mma_kk(kColDP, S::kV, S::kdOA);
        commit(dp_full);
        for (int it = 0; it < n_it; ++it) {
          const int s = it & 1;
          const uint32_t acc = it > 0 ? 1u : 0u;
          // dV += P^T dO_i  (B: 128 q x 64 d of each CTA, MN-major)
          mbar_wait_cluster(p_full, it & 1);
          tc_fence_after();
          BWD_STAMP(it, 0);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              umma2_ts_lo(tm + kColDV, tm + kColS + (ks >> 2) * 64 + (ks & 3) * 8, m_lo,
                          (S::kdOB + ks * 2048) >> 4, m_hi, id2_tv, (acc | (ks > 0)) ? 1u : 0u);
          }
          commit(do_empty);
          if (it + 1 < n_it) {
            mbar_wait_cluster(&q_full[s ^ 1], ((it + 1) >> 1) & 1);
            tc_fence_after();
            mma_kk(kColS, S::kK, S::kQA + static_cast<uint32_t>((s ^ 1) * S::kHalf));
            commit(s_full);
          }
          // dQ_i (this CTA's share)
          mbar_wait(ds_local, it & 1);
          tc_fence_after();
          dq_local();
          // dK += dS^T Q_i  (A: dS^T K-major; B: 128 q x 64 d of each CTA, MN-major)
          mbar_wait_cluster(ds_pair, it & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t off = (ks >> 2) * S::kBox + (ks & 3) * 32;
              umma2_ss_lo(tm + kColDK, k_lo, (S::kdS + off) >> 4, k_hi, m_lo,
                          (S::kQB + s * S::kHalf + ks * 2048) >> 4, m_hi, id2_tv, (acc | (ks > 0)) ? 1u : 0u);
            }
          }
          commit(&q_empty[s]);
          // dP^T(i+1): its TMEM columns hold dQ_i of BOTH CTAs until their drain warpgroups read them
          if (it + 1 < n_it) {
            mbar_wait_cluster(do_full, (it + 1) & 1);
            mbar_wait_cluster(dq_empty, it & 1);
            tc_fence_after();
            mma_kk(kColDP, S::kV, S::kdOA);
            commit(dp_full);
          }
        }
        commit(dkdv_full);
      } else {
        for (int it = 0; it < n_it; ++it) {
          mbar_wait(ds_local, it & 1);
          tc_fence_after();
          dq_local();
        }
      }
    }
  } else if (warp < 12) {
    // ================================ compute warpgroups ===============================
    setmaxnreg_inc<136>();
    const int half = (warp - 4) >> 2;  // which 64 q columns
    const int wq = warp & 3;
    const int row = wq * 32 + lane;    // key row within the block
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const int c0 = half * 64;
    const float sl2 = p.scale_log2;
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      const int i = i0 + it % nqi;
      // ---- P^T ----
      mbar_wait(&stat_full[s], (it >> 1) & 1);
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      uint32_t sr[2][32];
      tmem_ld_x32(tmem_base + lane_off + kColS + c0, sr[0]);
      tmem_ld_x32(tmem_base + lane_off + kColS + c0 + 32, sr[1]);
      tmem_ld_wait();
      float pf[64];
      const float4* l4 = reinterpret_cast<const float4*>(s_lse + s * 128 + c0);
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float4 l = l4[u];
        pf[4 * u + 0] = fast_exp2(fmaf(__uint_as_float(sr[u >> 3][(4 * u + 0) & 31]), sl2, -l.x));
        pf[4 * u + 1] = fast_exp2(fmaf(__uint_as_float(sr[u >> 3][(4 * u + 1) & 31]), sl2, -l.y));
        pf[4 * u + 2] = fast_exp2(fmaf(__uint_as_float(sr[u >> 3][(4 * u + 2) & 31]), sl2, -l.z));
        pf[4 * u + 3] = fast_exp2(fmaf(__uint_as_float(sr[u >> 3][(4 * u + 3) & 31]), sl2, -l.w));
      }
      if (p.causal && i == j) {  // diagonal block: key k0+row is visible to query q0+c iff row <= c
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (row > c0 + c) pf[c] = 0.f;
      }
      if ((p.causal && i < j) || !key_keep) {  // the pair's first q block lies above CTA 1's diagonal
#pragma unroll
        for (int c = 0; c < 64; ++c) pf[c] = 0.f;
      }
      {
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) pk[c] = pack2<T>(pf[2 * c], pf[2 * c + 1]);
        tmem_st_x32(tmem_base + lane_off + kColS + half * 64, pk);  // inside this half's own S^T columns
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(p_full, 0);
      // ---- dS^T ----
      mbar_wait(dp_full, it & 1);
      tc_fence_after();
      tmem_ld_x32(tmem_base + lane_off + kColDP + c0, sr[0]);
      tmem_ld_x32(tmem_base + lane_off + kColDP + c0 + 32, sr[1]);
      tmem_ld_wait();
      const float4* d4 = reinterpret_cast<const float4*>(s_del + s * 128 + c0);
      uint8_t* drow = sdS + half * S::kBox + row * 128;
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
        uint4 v;
        v.x = pack2<T>(ds[0], ds[1]);
        v.y = pack2<T>(ds[2], ds[3]);
        v.z = pack2<T>(ds[4], ds[5]);
        v.w = pack2<T>(ds[6], ds[7]);
        *reinterpret_cast<uint4*>(drow + ((ch ^ (row & 7)) << 4)) = v;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(ds_local);
        mbar_arrive_cluster(ds_pair, 0);
      }
    }
    // ---- epilogue: dV (half 0) / dK (half 1) -> 16-bit -> swizzled smem -> TMA store ------
    if (n_it > 0) {
      mbar_wait(dkdv_full, 0);
      tc_fence_after();
    }
    {
      const uint32_t tsrc = tmem_base + lane_off + (half ? kColDK : kColDV);
      const float mul = half ? p.scale : 1.f;
      uint8_t* stage = (half ? sQB : sQA);  // 32 KB each: the Q stages are free by now
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t r[32];
        if (n_it > 0) {
          tmem_ld_x32(tsrc + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int x = 0; x < 32; ++x) r[x] = 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 v;
          v.x = pack2<T>(__uint_as_float(r[8 * u + 0]) * mul, __uint_as_float(r[8 * u + 1]) * mul);
          v.y = pack2<T>(__uint_as_float(r[8 * u + 2]) * mul, __uint_as_float(r[8 * u + 3]) * mul);
          v.z = pack2<T>(__uint_as_float(r[8 * u + 4]) * mul, __uint_as_float(r[8 * u + 5]) * mul);
          v.w = pack2<T>(__uint_as_float(r[8 * u + 6]) * mul, __uint_as_float(r[8 * u + 7]) * mul);
          const int chunk = c * 4 + u;
          const int bx = chunk >> 3, cin = chunk & 7;
          *reinterpret_cast<uint4*>(stage + bx * S::kBox + row * 128 + ((cin ^ (row & 7)) << 4)) = v;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + half, 128);
      if (wq == 0 && lane == 0) {
#pragma unroll
        for (int bx = 0; bx < 2; ++bx)
          tma_store_3d(half ? &tm_dk : &tm_dv, stage + bx * S::kBox, bx * 64, k0, bh_kv);
        bulk_commit();
        bulk_wait_read<0>();
      }
    }
  } else {
    // ================================ dQ drain warpgroup ===============================
    setmaxnreg_inc<152>();
    const int wq = warp & 3;
    const int row = wq * 32 + lane;  // query row within the block
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const bool issuer = (warp == 12 && lane == 0);
    int nred = 0;
    for (int it = 0; it < n_it; ++it) {
      const int bh_q = b * p.QH + hk * g + it / nqi;
      const int q0 = (i0 + it % nqi) * 128;
      mbar_wait(dq_full, it & 1);
      tc_fence_after();
      uint32_t r[D / 32][32];
#pragma unroll
      for (int c = 0; c < D / 32; ++c) tmem_ld_x32(tmem_base + lane_off + kColDP + c * 32, r[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(dq_empty, 0);
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint8_t* stage = sdQ + (nred & 1) * 16384;
        if (issuer) bulk_wait_read<1>();
        named_bar_sync(3, 128);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint4 v = make_uint4(r[c][4 * u], r[c][4 * u + 1], r[c][4 * u + 2], r[c][4 * u + 3]);
          *reinterpret_cast<uint4*>(stage + row * 128 + ((u ^ (row & 7)) << 4)) = v;
        }
        fence_proxy_async_smem();
        named_bar_sync(3, 128);
        if (issuer) {
          tma_reduce_add_3d(&tm_dqa, stage, c * 32, q0, bh_q);
          bulk_commit();
        }
        ++nred;
      }
    }
    if (issuer) bulk_wait<0>();
  }

  // ---- teardown: both CTAs must be done with the pair's TMEM before it is released ---------
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------
// prep: delta, lse2 (padded), zero dQ accumulator.  LPR lanes per row, one 16-byte vector each.
// ---------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(float* __restrict__ deltap, float* __restrict__ lse2p,
                     float* __restrict__ dq_accum, const T* __restrict__ dO,
                     const T* __restrict__ o, const float* __restrict__ lse, int QL, int QLp,
                     int64_t n_rows_p, int* __restrict__ tile_counter) {
  constexpr int LPR = D / 8;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (gid == 0) *tile_counter = 0;  // tile queue of the persistent main kernel
  const int64_t rowp = gid / LPR;  // padded row index: bh * QLp + q
  const int li = static_cast<int>(gid % LPR);
  if (rowp >= n_rows_p) return;  // LPR divides 32 and 256: whole row groups exit together
  const int64_t bh = rowp / QLp;
  const int q = static_cast<int>(rowp % QLp);
  float acc = 0.f;
  if (q < QL) {
    const int64_t off = (bh * QL + q) * D + li * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(dO + off);
    const uint4 c = *reinterpret_cast<const uint4*>(o + off);
    const T* ah = reinterpret_cast<const T*>(&a);
    const T* ch = reinterpret_cast<const T*>(&c);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc = fmaf(to_f32<T>(ah[e]), to_f32<T>(ch[e]), acc);
    float4* z = reinterpret_cast<float4*>(dq_accum + off);
    z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int sft = 1; sft < LPR; sft <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
  if (li == 0) {
    deltap[rowp] = q < QL ? acc : 0.f;
    // a fully masked row has lse = -inf: give it +inf like the padding rows so that P = 0
    const float l = q < QL ? lse[bh * QL + q] : INFINITY;
    lse2p[rowp] = l == -INFINITY ? INFINITY : l * kLog2e;
  }
}

// prep for packed sequences: grid (128-row blocks of the padded statistics, QH).  Statistics block s belongs to the
// sequence z with the largest packed_stat_row(cu_q[z], z) <= 128 s (every sequence is padded to whole blocks, so a
// block never straddles two; blocks in the gaps between sequences are never read).  (r02a launched (row groups of
// the LONGEST sequence, QH, nseq) CTAs: on config C4 -- 64 lengths between 128 and 16 384 -- 1.7 of 2.1 million CTAs
// found nothing to do, 0.9 ms of a 31 ms backward.)
template <typename T, int D>
__global__ void __launch_bounds__(256)
attn_bwd_prep_packed_kernel(float* __restrict__ deltap, float* __restrict__ lse2p,
                            float* __restrict__ dq_accum, const T* __restrict__ dO,
                            const T* __restrict__ o, const float* __restrict__ lse,
                            const int* __restrict__ cu_q, int nseq, int64_t total_q, int64_t QLp) {
  constexpr int LPR = D / 8;
  constexpr int kRows = 256 / LPR;
  const int sblk = blockIdx.x, h = blockIdx.y;
  int lo = 0, hi = nseq;   // largest z with (cu_q[z] >> 7) + z <= sblk
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((cu_q[mid] >> 7) + mid <= sblk) lo = mid; else hi = mid;
  }
  const int z = lo;
  const int q_off = cu_q[z];
  const int QL = cu_q[z + 1] - q_off;
  const int r0 = (sblk - ((q_off >> 7) + z)) << 7;   // first row of this block within the sequence
  if (r0 >= ((QL + 127) & ~127)) return;             // gap block (whole CTA)
  const int li = threadIdx.x % LPR;
#pragma unroll 2
  for (int it = 0; it < 128 / kRows; ++it) {
    const int r = r0 + it * kRows + threadIdx.x / LPR;
    float acc = 0.f;
    if (r < QL) {
      const int64_t off = (static_cast<int64_t>(h) * total_q + q_off + r) * D + li * 8;
      const uint4 a = *reinterpret_cast<const uint4*>(dO + off);
      const uint4 c = *reinterpret_cast<const uint4*>(o + off);
      const T* ah = reinterpret_cast<const T*>(&a);
      const T* ch = reinterpret_cast<const T*>(&c);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(to_f32<T>(ah[e]), to_f32<T>(ch[e]), acc);
      float4* zp = reinterpret_cast<float4*>(dq_accum + off);
      zp[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      zp[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int sft = 1; sft < LPR; sft <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
    if (li == 0) {
      const int64_t pr = static_cast<int64_t>(h) * QLp + (static_cast<int64_t>(sblk) << 7) + (r - r0);
      deltap[pr] = r < QL ? acc : 0.f;
      const float l = r < QL ? lse[static_cast<int64_t>(h) * total_q + q_off + r] : INFINITY;
      lse2p[pr] = l == -INFINITY ? INFINITY : l * kLog2e;
    }
  }
}

// post: dQ = T(scale * dQ_accum), 8 elements per thread
template <typename T>
__global__ void __launch_bounds__(256)
attn_bwd_post_kernel(T* __restrict__ dq, const float* __restrict__ dq_accum, int64_t n8,
                     float scale) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n8) return;
  const float4 a = reinterpret_cast<const float4*>(dq_accum)[2 * i];
  const float4 c = reinterpret_cast<const float4*>(dq_accum)[2 * i + 1];
  uint4 v;
  v.x = pack2<T>(a.x * scale, a.y * scale);
  v.y = pack2<T>(a.z * scale, a.w * scale);
  v.z = pack2<T>(c.x * scale, c.y * scale);
  v.w = pack2<T>(c.z * scale, c.w * scale);
  reinterpret_cast<uint4*>(dq)[i] = v;
}

// packed mode, persistent kernel: tile_pre[z] = sum over sequences before z of ceil(KL_z / 128) * KH
// (one block; each thread scans a contiguous chunk of sequences), and the tile counter is reset
__global__ void __launch_bounds__(256)
packed_tile_prefix_kernel(int* __restrict__ tile_pre, int* __restrict__ tile_counter,
                          const int* __restrict__ cu_k, int nseq, int KH) {
  __shared__ int part[256];
  const int per = (nseq + 255) / 256;
  const int z0 = min(nseq, static_cast<int>(threadIdx.x) * per), z1 = min(nseq, z0 + per);
  int sum = 0;
  for (int z = z0; z < z1; ++z) sum += ((cu_k[z + 1] - cu_k[z] + 127) >> 7) * KH;
  part[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < 256; ++i) { const int v = part[i]; part[i] = run; run += v; }
    tile_pre[nseq] = run;
    *tile_counter = 0;
  }
  __syncthreads();
  int run = part[threadIdx.x];
  for (int z = z0; z < z1; ++z) {
    tile_pre[z] = run;
    run += ((cu_k[z + 1] - cu_k[z] + 127) >> 7) * KH;
  }
}

inline size_t align256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

// rows of the padded per-row statistics: dense (QL rounded up to 128) or packed (every sequence
// padded to whole 128-row blocks: at most total_q / 128 + nseq blocks)
inline int64_t stat_rows(int QL, int64_t total_q, int nseq, bool packed) {
  return packed ? (total_q / 128 + nseq) * 128 : static_cast<int64_t>((QL + 127) / 128) * 128;
}

template <typename T, int D>
int launch_bwd(const AttnParams& a) {
  using S = BwdSmem<D>;
  const bool packed = a.cu_q != nullptr;
  const int64_t QLp = stat_rows(a.QL, a.total_q, a.nseq, packed);
  const int64_t BH = packed ? a.QH : static_cast<int64_t>(a.B) * a.QH;
  const int64_t rows_q = packed ? a.total_q : a.QL;
  const int64_t rows_k = packed ? a.total_k : a.KL;
  // workspace carve-up (a.delta is the workspace base)
  char* ws = reinterpret_cast<char*>(a.delta);
  const size_t stat_bytes = align256(static_cast<size_t>(BH) * QLp * sizeof(float));
  float* deltap = reinterpret_cast<float*>(ws);
  float* lse2p = reinterpret_cast<float*>(ws + stat_bytes);
  float* dqa = reinterpret_cast<float*>(ws + 2 * stat_bytes);
  // the persistent kernel's tile counter (and, packed, the per-sequence tile prefix) sit behind the
  // dQ accumulator
  // (a.E < D: embedding dims 16 / 32 run the D = 64 kernels -- the TMA maps below describe the real E-wide
  // rows and the boxes stay 64 wide, so columns >= E arrive as zeros in shared memory and are clipped on
  // the way out; every plain-pointer access uses the real E)
  const int E = a.E;
  int* tile_counter = reinterpret_cast<int*>(
      ws + 2 * stat_bytes + align256(static_cast<size_t>(BH) * rows_q * E * sizeof(float)));
  int* tile_pre = tile_counter + 64;

  if (packed) {
    dim3 pg(static_cast<unsigned>(QLp / 128), a.QH);
    attn_bwd_prep_packed_kernel<T, D><<<pg, 256, 0, a.stream>>>(
        deltap, lse2p, dqa, static_cast<const T*>(a.dO), static_cast<const T*>(a.o), a.lse, a.cu_q, a.nseq,
        a.total_q, QLp);
    NNOP_LAUNCH_CHECK();
  } else {
    const int64_t n_rows_p = BH * QLp;
    const int64_t threads = n_rows_p * (E / 8);
    auto prep = E == D ? attn_bwd_prep_kernel<T, D> : (E == 32 ? attn_bwd_prep_kernel<T, 32> : attn_bwd_prep_kernel<T, 16>);
    prep<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, a.stream>>>(
        deltap, lse2p, dqa, static_cast<const T*>(a.dO), static_cast<const T*>(a.o), a.lse, a.QL,
        static_cast<int>(QLp), n_rows_p, tile_counter);
    NNOP_LAUNCH_CHECK();
  }
  alignas(64) CUtensorMap tq, tk, tv, tdo, tdk, tdv, tdqa;
  const uint64_t bhq = static_cast<uint64_t>(BH);
  const uint64_t bhk = packed ? a.KH : static_cast<uint64_t>(a.B) * a.KH;
  if (int rc = make_tmap_3d(&tq, a.q, a.dtype, E, rows_q, bhq, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tk, a.k, a.dtype, E, rows_k, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tv, a.v, a.dtype, E, rows_k, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tdo, a.dO, a.dtype, E, rows_q, bhq, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tdk, a.dk, a.dtype, E, rows_k, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tdv, a.dv, a.dtype, E, rows_k, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tdqa, dqa, NNOP_F32, E, rows_q, bhq, 32, 128)) return rc;
  BwdParams bp;
  bp.lse2p = lse2p; bp.deltap = deltap;
  bp.QL = a.QL; bp.KL = a.KL; bp.QH = a.QH; bp.KH = a.KH; bp.QLp = static_cast<int>(QLp);
  bp.causal = a.causal;
  bp.scale = a.scale; bp.scale_log2 = a.scale * kLog2e;
  bp.cu_q = a.cu_q; bp.cu_k = a.cu_k; bp.dk_ptr = a.dk; bp.dv_ptr = a.dv; bp.total_k = a.total_k;
  bp.kpad = packed ? nullptr : a.kpad;
  bp.pair_t = a.pair_t; bp.dpair_t = a.dpair_t; bp.KLp = a.KLp;
  bp.tile_pre = tile_pre; bp.nseq = a.nseq;
  bp.lpt_group = 0;
  if (!packed && a.causal) {
    // a few (batch, kv head) units at a time: their Q + dO streams (all q heads of the group) must share
    // L2 with the dQ accumulator traffic.  Measured on C2 / C3 (profiles/r01e_perf_lpt_order.txt): 4 units
    // +1.5 % / +5.4 %, 12 units 0 / +3.7 %, 32 units -2.7 % / -1 % against unit-by-unit order.
    const double qdo_bytes = 2.0 * a.QL * D * sizeof(T) * (static_cast<double>(a.QH) / a.KH);
    const int env = getenv("NNOP_BWD_LPT_GROUP") ? atoi(getenv("NNOP_BWD_LPT_GROUP")) : -1;
    int grp = static_cast<int>(16.0 * 1024 * 1024 / (qdo_bytes > 1 ? qdo_bytes : 1));
    grp = grp < 4 ? 4 : (grp > 16 ? 16 : grp);
    bp.lpt_group = env >= 0 ? env : grp;
    if (bp.lpt_group > a.KH * a.B) bp.lpt_group = a.KH * a.B;
  }
  const bool bias = a.pair != nullptr;
  if (bias && !a.pair_t_ready)
    if (int rc = attn_pair_to_head_major(a)) return rc;
  const int nkv = (a.KL + 127) / 128;
  // kernel variant (nnop_set_bwd_pair_mode / NNOP_BWD_PAIR): 0 = automatic (persistent kernel when the
  // tile queue is at least two rounds deep, else one CTA per tile), 1 = CTA pairs (experiment),
  // 2 = one CTA per tile, 3 = persistent, 100+n = persistent on n CTAs (tests)
  const int mode = bwd_pair_mode();
  bool use_pair = false;
  if constexpr (D == 128) use_pair = !packed && nkv >= 2 && mode == 1 && !bias;
  const int nq_blocks = (a.QL + 127) / 128;
  // packed: the exact tile count is only known on the device; this is its upper bound
  const int64_t n_tiles = packed ? (a.total_k / 128 + a.nseq) * a.KH : static_cast<int64_t>(nkv) * a.KH * a.B;
  const int num_sms = sm_count();
  const bool persist_ok = !bias && a.kpad == nullptr && (packed || !a.causal || nq_blocks >= nkv) &&
                          n_tiles < (1LL << 30);
  // 4 = persistent CTA pairs that exchange dQ halves over distributed shared memory (one L2 reduce-add of
  // half the width per CTA and step); dense problems with at least two kv blocks
  const bool use_duo = !use_pair && persist_ok && !packed && nkv >= 2 && mode == 4;
  const bool use_persist = !use_pair && !use_duo && persist_ok &&
                           (mode == 3 || mode >= 100 || (mode == 0 && n_tiles >= 2LL * num_sms));
  if (use_duo) {
    auto kern = attn_bwd_sm100_persist_kernel<T, D, true>;
    NNOP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    const int npairs = (nkv + 1) / 2;
    const int64_t n_pair_tiles = static_cast<int64_t>(npairs) * a.KH * a.B;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(kBwdThreads); cfg.dynamicSmemBytes = S::kTotal; cfg.stream = a.stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(2 * (num_sms / 2));
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
      (void)cudaGetLastError();
      max_clusters = num_sms / 2;
    }
    const int64_t clusters = n_pair_tiles < max_clusters ? n_pair_tiles : max_clusters;
    cfg.gridDim = dim3(static_cast<unsigned>(2 * clusters));
    const int n_t = static_cast<int>(n_pair_tiles);
    timing_begin(1, a.stream);
    NNOP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tq, tk, tv, tdo, tdk, tdv, tdqa, bp, tile_counter, n_t, npairs));
    timing_end(1, a.stream);
    NNOP_LAUNCH_CHECK();
  } else if (use_pair) {
    if constexpr (D == 128) {
      alignas(64) CUtensorMap tq64, tdo64;
      if (int rc = make_tmap_3d(&tq64, a.q, a.dtype, D, rows_q, bhq, 64, 64)) return rc;
      if (int rc = make_tmap_3d(&tdo64, a.dO, a.dtype, D, rows_q, bhq, 64, 64)) return rc;
      auto kern = attn_bwd_pair_kernel<T>;
      NNOP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PairSmem::kTotal));
      dim3 grid(2 * ((nkv + 1) / 2), a.KH, a.B);   // __cluster_dims__(2,1,1): blocks (2m, 2m+1) pair up
      timing_begin(1, a.stream);
      kern<<<grid, kBwdThreads, PairSmem::kTotal, a.stream>>>(tq, tq64, tk, tv, tdo, tdo64, tdk, tdv, tdqa, bp);
      timing_end(1, a.stream);
      NNOP_LAUNCH_CHECK();
    }
  } else if (use_persist) {
    auto kern = attn_bwd_sm100_persist_kernel<T, D>;
    NNOP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    const int ctas = mode >= 100 ? (mode - 100 < 1 ? 1 : mode - 100) : num_sms;
    const int grid = static_cast<int>(n_tiles < ctas ? n_tiles : ctas);
    if (packed) {
      packed_tile_prefix_kernel<<<1, 256, 0, a.stream>>>(tile_pre, tile_counter, a.cu_k, a.nseq, a.KH);
      NNOP_LAUNCH_CHECK();
    }
    timing_begin(1, a.stream);
    kern<<<grid, kBwdThreads, S::kTotal, a.stream>>>(tq, tk, tv, tdo, tdk, tdv, tdqa, bp, tile_counter,
                                                     static_cast<int>(n_tiles), nkv);
    timing_end(1, a.stream);
    NNOP_LAUNCH_CHECK();
  } else {
    auto kern = bias ? attn_bwd_sm100_kernel<T, D, true> : attn_bwd_sm100_kernel<T, D, false>;
    NNOP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    dim3 grid(nkv, a.KH, packed ? a.nseq : a.B);
    timing_begin(1, a.stream);
    kern<<<grid, kBwdThreads, S::kTotal, a.stream>>>(tq, tk, tv, tdo, tdk, tdv, tdqa, bp);
    timing_end(1, a.stream);
    NNOP_LAUNCH_CHECK();
  }
  {
    const int64_t n8 = BH * rows_q * E / 8;
    attn_bwd_post_kernel<T><<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, a.stream>>>(
        static_cast<T*>(a.dq), dqa, n8, a.scale);
    NNOP_LAUNCH_CHECK();
  }
  if (bias)
    if (int rc = attn_dpair_from_head_major(a)) return rc;
  return NNOP_OK;
}

}  // namespace

#ifdef NNOP_BWD_TRACE
extern "C" int nnop_debug_bwd_trace(long long* host_out, int n) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, g_bwd_trace, sizeof(long long) * n));
}
extern "C" int nnop_debug_bwd_cta_log(long long* host_out, int max_ctas, int reset) {
  int n = 0;
  cudaMemcpyFromSymbol(&n, g_bwd_cta_n, sizeof(int));
  if (n > max_ctas) n = max_ctas;
  if (n > 1024) n = 1024;
  cudaMemcpyFromSymbol(host_out, g_bwd_cta_log, sizeof(long long) * 16 * n);
  if (reset) { int z = 0; cudaMemcpyToSymbol(g_bwd_cta_n, &z, sizeof(int)); }
  return n;
}
#endif

bool attn_sm100_bwd_available() { return true; }
void attn_sm100_set_bwd_pair_mode(int mode) { g_bwd_pair_mode.store(mode); }

size_t attn_sm100_bwd_workspace_bytes(int E, int QL, int QH, int B) {
  const size_t QLp = static_cast<size_t>((QL + 127) / 128) * 128;
  const size_t BH = static_cast<size_t>(B) * QH;
  return 2 * align256(BH * QLp * sizeof(float)) + align256(BH * static_cast<size_t>(QL) * E * sizeof(float)) + 256;
}

size_t attn_sm100_bwd_packed_workspace_bytes(int E, int64_t total_q, int nseq, int QH) {
  const size_t QLp = static_cast<size_t>(stat_rows(0, total_q, nseq, true));
  return 2 * align256(static_cast<size_t>(QH) * QLp * sizeof(float)) +
         align256(static_cast<size_t>(QH) * static_cast<size_t>(total_q) * E * sizeof(float)) +
         align256((static_cast<size_t>(nseq) + 2 + 64) * sizeof(int));
}

int attn_sm100_bwd(const AttnParams& a) {
  // E = 128 -> the D = 128 kernels; E in {16, 32, 64} -> the D = 64 kernels (narrower rows are zero-padded by TMA)
  if (a.dtype == NNOP_BF16)
    return a.E == 128 ? launch_bwd<__nv_bfloat16, 128>(a) : launch_bwd<__nv_bfloat16, 64>(a);
  return a.E == 128 ? launch_bwd<__half, 128>(a) : launch_bwd<__half, 64>(a);
}

}  // namespace nnop
