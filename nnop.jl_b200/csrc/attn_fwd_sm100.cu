// attn_fwd_sm100.cu -- flash attention forward on tcgen05 / TMEM / TMA (bf16 / fp16, E in
// {64,128}, causal or not, GQA, ragged QL/KL).  Replaces `_flash_attention_fwd!`
// (src/attention.jl:1-131), which runs 32x32 SIMT tiles with scalar FMAs.
//
// One CTA = 256 query rows (two 128-row tiles, "ping-pong") of one (q-head, batch) pair.
//   warp 0        TMA producer: Q tiles once, then K_0,V_0,K_1,V_1,... through a 4-slot ring
//   warp 1        MMA issuer (one thread): S_t = Q_t K_i^T (smem x smem, both K-major) into
//                 TMEM, O_t += P_t V_i (A = P_t in TMEM, B = V_i in smem, MN-major)
//   warp 2        TMEM allocator (512 columns: S_0, S_1 at 0/128, O_0, O_1 at 256/256+E)
//   warps 4-7     softmax warpgroup of tile 0: thread <-> query row (TMEM lane), so row max
//   warps 8-11    and row sum need no shuffles.  P (16-bit) overwrites S in TMEM.
// While one tile's warpgroup exponentiates, the tensor core works on the other tile.
// O stays un-normalised in TMEM; it is rescaled lazily (only when the running max grows by
// more than 2^8, exact because the same reference max is used for P, l and O) and divided by
// l once in the epilogue, which also emits lse = m + log(l) (one fp32 residual instead of the
// reference's (ms, ls), src/attention.jl:166-168).  Causal: kv blocks past the diagonal are
// never loaded (src/attention.jl:47), only the diagonal block is masked element-wise; CTAs
// are launched heaviest-first.
//
// Ragged sizes: TMA zero-fills rows past QL / KL (per head: the tensor maps are 3-D), key
// columns >= KL are masked to -inf, and the TMA store clips rows >= QL.
#include "common.cuh"
#include "internal.h"

namespace nnop {
namespace {

constexpr int kFwdThreads = 384;
constexpr int kNStage = 4;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units

struct FwdParams {
  float* lse;
  int QL, KL, QH, KH, causal;
  float scale_log2;
};

template <int D>
struct FwdSmem {
  static constexpr int kTileBytes = 128 * D * 2;   // one Q tile / K block / V block
  static constexpr int kBoxBytes = 128 * 64 * 2;   // one 64-column TMA box (16 KB)
  static constexpr int kNBox = D / 64;
  static constexpr int kQOff = 0;
  static constexpr int kKVOff = 2 * kTileBytes;
  static constexpr int kBarOff = kKVOff + kNStage * kTileBytes;
  static constexpr int kNumBars = 2 + 2 * kNStage + 6;
  static constexpr int kTotal = kBarOff + kNumBars * 8 + 16;
  static constexpr int kDynBytes = kTotal + 1024;  // slack for manual 1024-byte alignment
};

template <typename T, int D>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_q,
                      const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v,
                      const __grid_constant__ CUtensorMap tm_o, const FwdParams p) {
  using S = FwdSmem<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + S::kQOff;
  uint8_t* sKV = smem + S::kKVOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* q_full = bars;                     // [2]
  uint64_t* kv_full = bars + 2;                // [kNStage]
  uint64_t* kv_empty = bars + 2 + kNStage;     // [kNStage]
  uint64_t* s_full = bars + 2 + 2 * kNStage;   // [2]
  uint64_t* p_full = s_full + 2;               // [2]
  uint64_t* o_full = s_full + 4;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + S::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- work assignment ----------------------------------------------------------------
  const int qt = p.causal ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;  // heaviest first
  const int q0 = qt * 256;
  const int h = blockIdx.y, b = blockIdx.z;
  const int bh_q = b * p.QH + h;
  const int bh_kv = b * p.KH + h / (p.QH / p.KH);
  const bool act1 = q0 + 128 < p.QL;
  const int nb0 = ((p.causal ? min(p.KL, q0 + 128) : p.KL) + 127) >> 7;
  const int nb1 = act1 ? (((p.causal ? min(p.KL, q0 + 256) : p.KL) + 127) >> 7) : 0;
  const int nblk = act1 ? nb1 : nb0;

  // ---- one-time setup -----------------------------------------------------------------
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
    mbar_init(&q_full[0], 1);
    mbar_init(&q_full[1], 1);
    for (int i = 0; i < kNStage; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&p_full[t], 128);
      mbar_init(&o_full[t], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    setmaxnreg_dec<72>();  // releases 128*(168-72) = 12288 regs = 256*(216-168)
    if (warp == 0 && lane == 0) {
      // ================================ TMA producer =================================
      mbar_arrive_expect_tx(&q_full[0], S::kTileBytes);
#pragma unroll
      for (int bx = 0; bx < S::kNBox; ++bx)
        tma_load_3d(sQ + bx * S::kBoxBytes, &tm_q, &q_full[0], bx * 64, q0, bh_q);
      int n = 0;
      auto load_kv = [&](const CUtensorMap* tm, int blk) {
        const int st = n % kNStage;
        const uint32_t ph = (n / kNStage) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], S::kTileBytes);
#pragma unroll
        for (int bx = 0; bx < S::kNBox; ++bx)
          tma_load_3d(sKV + st * S::kTileBytes + bx * S::kBoxBytes, tm, &kv_full[st], bx * 64,
                      blk * 128, bh_kv);
        ++n;
      };
      load_kv(&tm_k, 0);
      if (act1) {
        mbar_arrive_expect_tx(&q_full[1], S::kTileBytes);
#pragma unroll
        for (int bx = 0; bx < S::kNBox; ++bx)
          tma_load_3d(sQ + S::kTileBytes + bx * S::kBoxBytes, &tm_q, &q_full[1], bx * 64, q0 + 128,
                      bh_q);
      }
      load_kv(&tm_v, 0);
      for (int i = 1; i < nblk; ++i) {
        load_kv(&tm_k, i);
        load_kv(&tm_v, i);
      }
    } else if (warp == 1 && lane == 0) {
      // ================================ MMA issuer ===================================
      constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, is_bf16<T>::value, false, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, D, is_bf16<T>::value, false, true);
      const uint32_t q_base = smem_u32(sQ);
      const uint32_t kv_base = smem_u32(sKV);
      auto slot_wait = [&](int slot) {
        mbar_wait(&kv_full[slot % kNStage], (slot / kNStage) & 1);
      };
      auto qk = [&](int t, int slot) {
        const uint32_t a0 = q_base + t * S::kTileBytes;
        const uint32_t b0 = kv_base + (slot % kNStage) * S::kTileBytes;
        const uint32_t d = tmem_base + t * 128;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const uint32_t off = (ks >> 2) * S::kBoxBytes + (ks & 3) * 32;
          umma_ss(d, make_smem_desc_sw128(a0 + off, 16, 1024),
                  make_smem_desc_sw128(b0 + off, 16, 1024), idesc_qk, ks > 0 ? 1u : 0u);
        }
      };
      auto pv = [&](int t, int slot, bool acc) {
        const uint32_t b0 = kv_base + (slot % kNStage) * S::kTileBytes;
        const uint32_t d = tmem_base + 256 + t * D;
        const uint32_t a = tmem_base + t * 128;  // P aliases S
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_ts(d, a + j * 8, make_smem_desc_sw128(b0 + j * 2048, S::kBoxBytes, 1024), idesc_pv,
                  (acc || j > 0) ? 1u : 0u);
      };
      mbar_wait(&q_full[0], 0);
      slot_wait(0);
      tc_fence_after();
      qk(0, 0);
      tc_commit(&s_full[0]);
      if (act1) {
        mbar_wait(&q_full[1], 0);
        tc_fence_after();
        qk(1, 0);
        tc_commit(&s_full[1]);
      }
      tc_commit(&kv_empty[0]);
      for (int i = 0; i < nblk; ++i) {
        const int vslot = 2 * i + 1, knext = 2 * i + 2;
        slot_wait(vslot);
        if (i < nb0) {
          mbar_wait(&p_full[0], i & 1);
          tc_fence_after();
          pv(0, vslot, i > 0);
          if (i + 1 < nb0) {
            slot_wait(knext);
            tc_fence_after();
            qk(0, knext);
            tc_commit(&s_full[0]);
          } else {
            tc_commit(&o_full[0]);
          }
        }
        if (act1) {
          mbar_wait(&p_full[1], i & 1);
          tc_fence_after();
          pv(1, vslot, i > 0);
          if (i + 1 < nblk) {
            slot_wait(knext);
            tc_fence_after();
            qk(1, knext);
            tc_commit(&s_full[1]);
          } else {
            tc_commit(&o_full[1]);
          }
        }
        tc_commit(&kv_empty[vslot % kNStage]);
        if (i + 1 < nblk) tc_commit(&kv_empty[knext % kNStage]);
      }
    }
  } else {
    // ================================ softmax warpgroups ===============================
    setmaxnreg_inc<216>();
    const int t = (warp - 4) >> 2;
    const int nbt = t ? nb1 : nb0;
    if (nbt > 0) {
      const int wq = warp & 3;
      const int row = wq * 32 + lane;
      const int q_row = q0 + t * 128 + row;
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tS = tmem_base + lane_off + t * 128;
      const uint32_t tO = tmem_base + lane_off + 256 + t * D;
      const float sl2 = p.scale_log2;
      float m_used = -INFINITY;  // reference max in scaled log2 units
      float l = 0.f;

      for (int i = 0; i < nbt; ++i) {
        mbar_wait(&s_full[t], i & 1);
        tc_fence_after();
        uint32_t sr[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, sr[c]);
        tmem_ld_wait();

        const int k0 = i * 128;
        const bool need_mask = (k0 + 128 > p.KL) || (p.causal && (k0 + 127 > q0 + t * 128));
        if (need_mask) {
          const int lim = p.causal ? min(p.KL - 1, q_row) : (p.KL - 1);  // last visible key
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (k0 + c * 32 + j > lim) sr[c][j] = 0xff800000u;  // -inf
        }
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(sr[c][j]));
        const float mx_s = mx * sl2;

        if (i == 0) {
          m_used = (mx_s == -INFINITY) ? 0.f : mx_s;
        } else {
          const bool grow = mx_s > m_used + kRescaleThreshold;
          if (__any_sync(0xffffffffu, grow)) {
            const float m_new = grow ? mx_s : m_used;
            const float alpha = fast_exp2(m_used - m_new);
            m_used = m_new;
            l *= alpha;
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
              uint32_t orow[32];
              tmem_ld_x32(tO + c * 32, orow);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) orow[j] = __float_as_uint(__uint_as_float(orow[j]) * alpha);
              tmem_st_x32(tO + c * 32, orow);
            }
          }
        }

        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pr[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(sr[c][2 * j]), sl2, -m_used));
            const float p1 = fast_exp2(fmaf(__uint_as_float(sr[c][2 * j + 1]), sl2, -m_used));
            sum += p0 + p1;
            pr[j] = pack2<T>(p0, p1);
          }
          tmem_st_x16(tS + c * 16, pr);
        }
        l += sum;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[t]);
      }

      // ---- epilogue: O / l -> 16-bit -> swizzled smem (the Q_t buffer) -> TMA store -----
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      const float inv_l = l > 0.f ? 1.f / l : 0.f;
      uint8_t* stage = sQ + t * S::kTileBytes;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t orow[32];
        tmem_ld_x32(tO + c * 32, orow);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 v;
          v.x = pack2<T>(__uint_as_float(orow[8 * u + 0]) * inv_l, __uint_as_float(orow[8 * u + 1]) * inv_l);
          v.y = pack2<T>(__uint_as_float(orow[8 * u + 2]) * inv_l, __uint_as_float(orow[8 * u + 3]) * inv_l);
          v.z = pack2<T>(__uint_as_float(orow[8 * u + 4]) * inv_l, __uint_as_float(orow[8 * u + 5]) * inv_l);
          v.w = pack2<T>(__uint_as_float(orow[8 * u + 6]) * inv_l, __uint_as_float(orow[8 * u + 7]) * inv_l);
          const int chunk = c * 4 + u;  // 16-byte chunk index within the row
          const int bx = chunk >> 3, cin = chunk & 7;
          *reinterpret_cast<uint4*>(stage + bx * S::kBoxBytes + row * 128 + ((cin ^ (row & 7)) << 4)) = v;
        }
      }
      if (q_row < p.QL)
        p.lse[static_cast<int64_t>(bh_q) * p.QL + q_row] =
            l > 0.f ? (m_used + fast_log2(l)) * kLn2 : -INFINITY;
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 128);
      if (wq == 0 && lane == 0) {
#pragma unroll
        for (int bx = 0; bx < S::kNBox; ++bx)
          tma_store_3d(&tm_o, stage + bx * S::kBoxBytes, bx * 64, q0 + t * 128, bh_q);
        bulk_commit();
        bulk_wait_read<0>();
      }
    }
  }

  // ---- teardown -------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <typename T, int D>
int launch_fwd(const AttnParams& a) {
  using S = FwdSmem<D>;
  alignas(64) CUtensorMap tq, tk, tv, to;
  const uint64_t bhq = static_cast<uint64_t>(a.B) * a.QH, bhk = static_cast<uint64_t>(a.B) * a.KH;
  if (int rc = make_tmap_3d(&tq, a.q, a.dtype, D, a.QL, bhq, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tk, a.k, a.dtype, D, a.KL, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tv, a.v, a.dtype, D, a.KL, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&to, a.o, a.dtype, D, a.QL, bhq, 64, 128)) return rc;
  auto kern = attn_fwd_sm100_kernel<T, D>;
  NNOP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kDynBytes));
  FwdParams fp;
  fp.lse = a.lse;
  fp.QL = a.QL; fp.KL = a.KL; fp.QH = a.QH; fp.KH = a.KH; fp.causal = a.causal;
  fp.scale_log2 = a.scale * kLog2e;
  dim3 grid((a.QL + 255) / 256, a.QH, a.B);
  timing_begin(0, a.stream);
  kern<<<grid, kFwdThreads, S::kDynBytes, a.stream>>>(tq, tk, tv, to, fp);
  timing_end(0, a.stream);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

}  // namespace

bool attn_sm100_supported(const AttnParams& a, bool backward) {
  if (backward && !attn_sm100_bwd_available()) return false;
  if (a.dtype != NNOP_F16 && a.dtype != NNOP_BF16) return false;
  if (a.E != 64 && a.E != 128) return false;
  if (a.pair || a.kpad) return false;
  if (a.QL < 1 || a.KL < 1) return false;
  if (a.QH > 65535 || a.B > 65535) return false;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(a.q) || !al(a.k) || !al(a.v) || !al(a.o)) return false;
  if (backward && (!al(a.dq) || !al(a.dk) || !al(a.dv) || !al(a.dO))) return false;
  return true;
}

int attn_sm100_fwd(const AttnParams& a) {
  if (a.dtype == NNOP_BF16)
    return a.E == 128 ? launch_fwd<__nv_bfloat16, 128>(a) : launch_fwd<__nv_bfloat16, 64>(a);
  return a.E == 128 ? launch_fwd<__half, 128>(a) : launch_fwd<__half, 64>(a);
}

}  // namespace nnop
