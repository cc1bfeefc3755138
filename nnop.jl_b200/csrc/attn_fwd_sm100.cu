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
//
// Variants in this file: attn_fwd_sm100_kernel<T, D, SPLIT, BIAS> (one CTA per 256-row q tile; SPLIT =
// Float32 operands as two fp16 terms, BIAS = additive pair bias streamed by TMA) and
// attn_fwd_sm100_persist_kernel<T, D> (one CTA per SM over a dynamic queue of q tiles; same softmax code).
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include "common.cuh"
#include "internal.h"

namespace nnop {
namespace {

constexpr int kFwdThreads = 384;
// QUAD variant: the producer warpgroup (TMA, MMA, TMEM allocator) and sixteen softmax warps (two per 32 rows of
// each tile); launched at 96 registers per thread, setmaxnreg moves them to 64 / 104
constexpr int kFwdQuadThreads = 640;
constexpr int kFwdQuadXchBytes = 6144;   // row-max exchange [2 tiles][2 parities][2 halves][128] + row-sum exchange [2][2][128]
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
#ifndef NNOP_FWD_POLY_EVERY
#define NNOP_FWD_POLY_EVERY 0
#endif
// every kPolyEvery-th pair of exponentials runs on the FMA pipe instead of the MUFU (0 = none)
constexpr int kPolyEvery = NNOP_FWD_POLY_EVERY;
#ifndef NNOP_FWD_QUAD_POLY_EVERY
#define NNOP_FWD_QUAD_POLY_EVERY 0
#endif
constexpr int kQuadPolyEvery = NNOP_FWD_QUAD_POLY_EVERY;   // QUAD variant: same switch
#ifndef NNOP_FWD_SPLIT_QK
#define NNOP_FWD_SPLIT_QK 0
#endif
#ifndef NNOP_FWD_SPLIT_PV
#define NNOP_FWD_SPLIT_PV 1
#endif
constexpr bool kSplitQK = NNOP_FWD_SPLIT_QK != 0;  // issue S columns 64.. of step i+1 before PV(i)
constexpr bool kSplitPV = NNOP_FWD_SPLIT_PV != 0;  // start PV on keys 0..63 while 64..127 exponentiate

#ifdef NNOP_FWD_TRACE
// development aid: pipeline timeline of CTA (0,0,0), 16 clock64 stamps per kv step
__device__ long long g_fwd_trace[256 * 16];
#define FWD_STAMP(i, k) do { if (tr) g_fwd_trace[(i) * 16 + (k)] = clock64(); } while (0)
// NNOP_FWD_TRACE=2: the four warps of tile 0's softmax warpgroup instead (warp w: stamps 4w .. 4w+3 = logits in
// registers, max decided, first / second half of P announced): how far apart do the warps of one group run?
#define FWD_TRACE_WARPS (NNOP_FWD_TRACE == 2)
#else
#define FWD_STAMP(i, k) do { } while (0)
#define FWD_TRACE_WARPS 0
#endif

struct FwdParams {
  float* lse;
  int QL, KL, QH, KH, causal;  // packed (varlen) mode: QL / KL are the maximum sequence lengths
  float scale_log2;
  const float* f32_mult;  // SPLIT (Float32) kernels: multipliers of the scale block (internal.h F32Mult), else unused
  // packed variable-length mode (cu_q != nullptr): sequence z = blockIdx.z owns rows
  // [cu_q[z], cu_q[z+1]) of the (QH, total_q, E) tensors and keys [cu_k[z], cu_k[z+1])
  const int* cu_q;
  const int* cu_k;
  void* o_ptr;       // raw output pointer for partial tiles (a TMA store would spill into the next sequence)
  int o_row_bytes;   // bytes of one output row as the caller sees it: E x sizeof(output element) (E may be < D)
  int64_t total_q;
  const uint8_t* kpad;  // (B, KL) key padding mask, 1 = attend, or nullptr (dense mode only)
  int nseq;             // packed mode: number of sequences
  // dense causal mode, one CTA per q tile: heads are walked in groups of `lpt_group` (batch, head) pairs
  // whose K / V fit in L2 together; inside a group all heaviest q tiles come first.  0 = head by head.
  int lpt_group;
};

// NS = stages of the K/V ring; BIASB = bytes of additive-bias staging (pair: 4 x 16 KB, ring of 3)
template <int D, int NS = 4, int BIASB = 0>
struct FwdSmem {
  static constexpr int kNStage = NS;
  static constexpr int kTileBytes = 128 * D * 2;   // one Q tile / K block / V block
  static constexpr int kBoxBytes = 128 * 64 * 2;   // one 64-column TMA box (16 KB)
  static constexpr int kNBox = D / 64;
  static constexpr int kQTiles = D == 256 ? 1 : 2;   // E = 256: one 128-row q tile per CTA (64 KB tiles)
  static constexpr int kQOff = 0;
  static constexpr int kKVOff = kQTiles * kTileBytes;
  static constexpr int kBiasOff = kKVOff + kNStage * kTileBytes;  // [2 tiles][2 buffers] x 16 KB
  static constexpr int kBarOff = kBiasOff + BIASB;
  static constexpr int kNumBars = 2 + 2 * kNStage + 10 + 8;
  static constexpr int kTotal = kBarOff + kNumBars * 8 + 16;
  static constexpr int kDynBytes = kTotal + 1024;  // slack for manual 1024-byte alignment
};

// SPLIT = true is the Float32 (E = 64) path: every fp32 operand x arrives as two fp16 terms
// [hi | lo] side by side in a 128-wide row (x ~ hi + lo, 22 significant bits; written by
// split_f32_kernel), S = Qh Kh^T + Qh Kl^T + Ql Kh^T is three chained MMAs, P is split in registers
// into Ph + Pl (two TMEM operands), O' = (Ph + Pl) [Vh | Vl] accumulates both V terms side by side
// and the epilogue adds the two halves and writes fp32.  Same tiles, barriers and shared-memory /
// TMEM footprint as the bf16 E = 128 kernel; 1.75x its tensor time for half its FLOPs.
//
// BIAS = true adds the `pair` term (src/attention.jl:55-62): S' = S * scale + pair[h, q, k, b].  The
// launcher first transposes pair into head-major (B, QH, QL, KLp) rows (attn_pair.cu) so that a
// (128 query) x (128-byte) box of it is one TMA load; warp 3 streams those boxes through two 16 KB
// buffers per tile (the K/V ring gives up one stage for them) and the softmax warpgroups fold them
// into the logits (in log2 units) before the mask / max / exp steps.
//
// QUAD = true (16-bit, E = 128, no bias): FOUR softmax warpgroups, two per tile, each thread owning HALF a row
// (64 key columns).  The per-tile chain S -> softmax -> P -> PV -> QK -> S bounds the step, and a single thread
// walking 128 logits is most of it (~2 100 of ~3 250 clk); two threads per row halve that.  The two
// warps that share 32 rows exchange their partial row max through shared memory behind a 64-thread named
// barrier (the speculative first chunk of exponentials is issued before the barrier, so its latency is
// hidden), keep partial row sums (added in the epilogue), rescale their own half of O's columns and write
// their P into the first 32 columns of their OWN half of S (so neither can overwrite logits the other has
// not read yet; the PV MMA takes its A operand from two 32-column pieces).
template <typename T, int D, bool SPLIT = false, bool BIAS = false, bool QUAD = false>
__global__ void __launch_bounds__(QUAD ? kFwdQuadThreads : kFwdThreads, 1)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tm_q,
                      const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v,
                      const __grid_constant__ CUtensorMap tm_o,
                      const __grid_constant__ CUtensorMap tm_bias, const FwdParams p) {
  // D = 256 (16-bit E = 256): ONE 128-row q tile per CTA, K / V blocks of 64 KB through a two-slot ring (192 KB of
  // shared memory), S at TMEM columns 0..127 and the 256-column O at 256..511; the second softmax warpgroup idles.
  // No ping-pong, so the tensor pipe waits for every softmax -- still tens of times the SIMT kernel.
  constexpr bool SINGLE = D == 256;
  // SINGLE + SPLIT = Float32 E = 128: rows [hi 128 | lo 128]
  static_assert(!SINGLE || (!BIAS && !QUAD), "256-wide rows: forward without a bias only");
  using S = FwdSmem<D, SINGLE ? 2 : (BIAS ? 3 : 4), BIAS ? 65536 : 0>;
  constexpr int kNStage = S::kNStage;
  constexpr int kCtaRows = SINGLE ? 128 : 256;
  static_assert(!QUAD || (D == 128 && !SPLIT && !BIAS), "QUAD: 16-bit E = 128 without a bias");
  constexpr int kThreads = QUAD ? kFwdQuadThreads : kFwdThreads;
  constexpr int kAllocWarp = 2;       // TMEM allocator
  constexpr int kFirstSoftmaxWarp = 4;
  using BT = typename std::conditional<SPLIT, float, T>::type;  // element type of the bias
  constexpr int kBiasCols = 128 / static_cast<int>(sizeof(BT));  // keys per 128-byte box row
  constexpr int kBiasChunks = 128 / kBiasCols;                   // boxes per 128-key block
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + S::kQOff;
  uint8_t* sKV = smem + S::kKVOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* q_full = bars;                     // [2]
  uint64_t* kv_full = bars + 2;                // [kNStage]
  uint64_t* kv_empty = bars + 2 + kNStage;     // [kNStage]
  uint64_t* s_full = bars + 2 + 2 * kNStage;   // [2]  S_t complete in TMEM
  uint64_t* s_read = s_full + 2;               // [2]  S_t now lives in registers (columns 64.. reusable)
  uint64_t* p_half = s_full + 4;               // [2][2]  P_t keys 0..63 / 64..127 written
  uint64_t* o_full = s_full + 8;               // [2]
  uint64_t* bias_full = s_full + 10;           // [2 tiles][2 buffers]
  uint64_t* bias_empty = s_full + 14;          // [2 tiles][2 buffers]
  uint8_t* sBias = smem + S::kBiasOff;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + S::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef NNOP_FWD_TRACE
  const bool tr = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
#endif

  // ---- work assignment ----------------------------------------------------------------
  const bool packed = p.cu_q != nullptr;
  int qt = p.causal ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;  // heaviest first
  int h = blockIdx.y;
  int QL = p.QL, KL = p.KL, q_off = 0, k_off = 0, b = blockIdx.z;
  if (!packed && p.lpt_group > 1) {
    // Blocks are dispatched in linear order (x fastest).  Head by head, the last head's heaviest tile
    // starts when the grid is ~99 % dispatched and runs alone for its whole length; walking the heads
    // in groups and dispatching each group heaviest-tile-first leaves only light tiles for the end,
    // while a group's K / V still share L2.
    const int nqt = gridDim.x, nbh = gridDim.y * gridDim.z;
    const int lin = blockIdx.x + nqt * (blockIdx.y + gridDim.y * blockIdx.z);
    const int per = p.lpt_group * nqt;
    const int g0 = (lin / per) * p.lpt_group;          // first (batch, head) pair of this group
    const int gsz = min(p.lpt_group, nbh - g0);        // the last group may be smaller
    const int idx = lin - (lin / per) * per;
    const int r = idx / gsz, u = idx - r * gsz;
    const int bh = g0 + u;
    qt = p.causal ? nqt - 1 - r : r;
    h = bh % gridDim.y;
    b = bh / gridDim.y;
  }
  if (packed) {
    // The grid has one CTA per 256-row q tile of the packed batch (upper bound total_q / 256 + nseq,
    // no CTAs for tiles a shorter sequence does not have).  Tile e = blockIdx.x belongs to the
    // sequence z whose tile range contains it: every thread counts the tiles of a chunk of
    // sequences, then all walk the per-thread counts (scratch: the not yet used Q buffer).
    int* scratch = reinterpret_cast<int*>(smem);
    const int e = blockIdx.x;
    const int per = (p.nseq + kThreads - 1) / kThreads;
    auto tiles_of = [&](int z) { return (p.cu_q[z + 1] - p.cu_q[z] + 255) >> 8; };
    {
      const int z0 = min(p.nseq, static_cast<int>(threadIdx.x) * per), z1 = min(p.nseq, z0 + per);
      int cnt = 0;
      for (int z = z0; z < z1; ++z) cnt += tiles_of(z);
      scratch[threadIdx.x] = cnt;
    }
    __syncthreads();
    int run = 0, c = 0;
    for (; c < kThreads; ++c) {
      const int v = scratch[c];
      if (e < run + v) break;
      run += v;
    }
    __syncthreads();   // scratch is the Q buffer from here on
    if (c == kThreads) return;  // past the last tile of the batch (whole CTA exits together)
    int z = c * per, nt = tiles_of(z);
    while (e >= run + nt) {
      run += nt;
      nt = tiles_of(++z);
    }
    qt = p.causal ? nt - 1 - (e - run) : e - run;  // heaviest tile of a sequence first
    q_off = p.cu_q[z];
    QL = p.cu_q[z + 1] - q_off;
    k_off = p.cu_k[z];
    KL = p.cu_k[z + 1] - k_off;
    b = 0;
  }
  const int q0 = qt * kCtaRows;
  const int bh_q = b * p.QH + h;
  const int bh_kv = b * p.KH + h / (p.QH / p.KH);
  // key padding mask (src/attention.jl:73-79): keys past the last attended one are never loaded
  // (a prefix-shaped mask costs nothing beyond its own length); blocks that still contain masked
  // keys are masked element-wise in the softmax warpgroups
  const uint8_t* kmask = p.kpad ? p.kpad + static_cast<int64_t>(b) * p.KL : nullptr;
  if (kmask) {
    __shared__ int s_kl_eff;
    if (threadIdx.x == 0) s_kl_eff = 0;
    __syncthreads();
    int last = 0;
    for (int kk = threadIdx.x; kk < KL; kk += kThreads)
      if (kmask[kk]) last = kk + 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, o));
    if (lane == 0 && last > 0) atomicMax(&s_kl_eff, last);
    __syncthreads();
    KL = s_kl_eff;
  }
  const bool act1 = !SINGLE && q0 + 128 < QL;
  const int nb0 = ((p.causal ? min(KL, q0 + 128) : KL) + 127) >> 7;
  const int nb1 = act1 ? (((p.causal ? min(KL, q0 + 256) : KL) + 127) >> 7) : 0;
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
      mbar_init(&s_read[t], 4);  // one arrival per softmax warp
      mbar_init(&p_half[2 * t], 4);
      mbar_init(&p_half[2 * t + 1], 4);
      mbar_init(&o_full[t], 1);
      for (int u = 0; u < 2; ++u) {
        mbar_init(&bias_full[2 * t + u], 1);
        mbar_init(&bias_empty[2 * t + u], 4);  // one arrival per softmax warp
      }
    }
    if constexpr (BIAS) tma_prefetch_desc(&tm_bias);
    fence_mbar_init();
  }
  if (warp == kAllocWarp) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kFirstSoftmaxWarp) {
    if constexpr (QUAD) setmaxnreg_dec<64>(); else setmaxnreg_dec<88>();  // launch: 384 x 168; warps 0-3 give back 128 x 80 = 10240 registers, exactly what 256 x (208 - 168) takes
    if (warp == 0 && lane == 0 && nblk > 0) {
      // ================================ TMA producer =================================
      mbar_arrive_expect_tx(&q_full[0], S::kTileBytes);
#pragma unroll
      for (int bx = 0; bx < S::kNBox; ++bx)
        tma_load_3d(sQ + bx * S::kBoxBytes, &tm_q, &q_full[0], bx * 64, q_off + q0, bh_q);
      int n = 0;
      auto load_kv = [&](const CUtensorMap* tm, int blk) {
        const int st = n % kNStage;
        const uint32_t ph = (n / kNStage) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], S::kTileBytes);
#pragma unroll
        for (int bx = 0; bx < S::kNBox; ++bx)
          tma_load_3d(sKV + st * S::kTileBytes + bx * S::kBoxBytes, tm, &kv_full[st], bx * 64,
                      k_off + blk * 128, bh_kv);
        ++n;
      };
      load_kv(&tm_k, 0);
      if (act1) {
        mbar_arrive_expect_tx(&q_full[1], S::kTileBytes);
#pragma unroll
        for (int bx = 0; bx < S::kNBox; ++bx)
          tma_load_3d(sQ + S::kTileBytes + bx * S::kBoxBytes, &tm_q, &q_full[1], bx * 64,
                      q_off + q0 + 128, bh_q);
      }
      load_kv(&tm_v, 0);
      if constexpr (SINGLE) {
        // K lives in slot 0 and V in slot 1 (uses counted per slot).  QK(i+1) completes before PV(i-1) does, so
        // K(i+2) can be fetched before V(i+1): issue the loads in the order their slots come free
        auto load_to = [&](const CUtensorMap* tm, int st, int blk) {
          mbar_wait(&kv_empty[st], (blk & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[st], S::kTileBytes);
#pragma unroll
          for (int bx = 0; bx < S::kNBox; ++bx)
            tma_load_3d(sKV + st * S::kTileBytes + bx * S::kBoxBytes, tm, &kv_full[st], bx * 64, k_off + blk * 128, bh_kv);
        };
        if (nblk > 1) load_to(&tm_k, 0, 1);
        for (int i = 1; i < nblk; ++i) {
          if (i + 1 < nblk) load_to(&tm_k, 0, i + 1);
          load_to(&tm_v, 1, i);
        }
      } else {
      for (int i = 1; i < nblk; ++i) {
        load_kv(&tm_k, i);
        load_kv(&tm_v, i);
      }
      }
    } else if (BIAS && warp == 3 && lane == 0 && nblk > 0) {
      // ================================ bias producer ================================
      // box c of block i of tile t -> buffer (i * kBiasChunks + c) & 1 of that tile, in the order the
      // softmax warpgroups consume them (the two tiles interleaved)
      for (int i = 0; i < nblk; ++i)
        for (int c = 0; c < kBiasChunks; ++c)
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (i >= (t ? nb1 : nb0)) continue;
            const int n = i * kBiasChunks + c;
            const int u = 2 * t + (n & 1);
            mbar_wait(&bias_empty[u], ((n >> 1) & 1) ^ 1);
            mbar_arrive_expect_tx(&bias_full[u], 16384);
            tma_load_3d(sBias + u * 16384, &tm_bias, &bias_full[u], i * 128 + c * kBiasCols, q0 + t * 128, bh_q);
          }
    } else if (warp == 1 && nblk > 0) {
      // ================================ MMA issuer ===================================
      // The whole warp runs the (uniform) control flow and the barrier waits; one elected lane
      // issues tcgen05.mma / commit.  Keeping every operand warp-uniform lets descriptors live in
      // uniform registers (a lane-0-only branch costs a local-memory reload + R2UR waterfall per MMA).
      constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, is_bf16<T>::value, false, false);
      constexpr uint32_t idesc_qk64 = make_idesc_f16(128, 64, is_bf16<T>::value, false, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, D, is_bf16<T>::value, false, true);
      const uint32_t tm = uniform_u32(tmem_base);
      const uint32_t q_base = uniform_u32(smem_u32(sQ));
      const uint32_t kv_base = uniform_u32(smem_u32(sKV));
      const uint64_t dq0 = make_smem_desc_sw128(q_base, 16, 1024);               // K-major operands
      const uint64_t dk0 = make_smem_desc_sw128(kv_base, 16, 1024);
      const uint64_t dv0 = make_smem_desc_sw128(kv_base, S::kBoxBytes, 1024);    // V as MN-major B
      auto slot_wait = [&](int slot) {
        mbar_wait(&kv_full[slot % kNStage], (slot / kNStage) & 1);
      };
      // descriptor address field is (byte address >> 4): advancing an operand = adding bytes/16.
      // (Building descriptors inside the asm block, as the backward does, measured 5% slower here.)
      // (SINGLE: `t` names the S buffer -- there is one q tile and the logits are double-buffered)
      auto qk = [&](int t, int slot) {
        const uint64_t a0 = dq0 + static_cast<uint64_t>(((SINGLE ? 0 : t) * S::kTileBytes) >> 4);
        const uint64_t b0 = dk0 + static_cast<uint64_t>(((slot % kNStage) * S::kTileBytes) >> 4);
        const uint32_t d = tm + t * 128;
        if (elect_one()) {
          if constexpr (SPLIT) {
            // first half of the boxes = hi terms, second half = lo terms: hi*hi, hi*lo, lo*hi (lo*lo is below fp32
            // rounding); D / 32 k-steps of 16 per part (E = D / 2)
            constexpr int kLo = (S::kNBox / 2) * S::kBoxBytes;
#pragma unroll
            for (int part = 0; part < 3; ++part)
#pragma unroll
              for (int k4 = 0; k4 < D / 32; ++k4) {
                const uint32_t koff = (k4 >> 2) * S::kBoxBytes + (k4 & 3) * 32;
                const uint32_t aoff = ((part == 2 ? kLo : 0) + koff) >> 4;
                const uint32_t boff = ((part == 1 ? kLo : 0) + koff) >> 4;
                umma_ss(d, a0 + aoff, b0 + boff, idesc_qk, (part | k4) ? 1u : 0u);
              }
          } else {
#pragma unroll
            for (int ks = 0; ks < D / 16; ++ks) {
              const uint32_t off = ((ks >> 2) * S::kBoxBytes + (ks & 3) * 32) >> 4;
              umma_ss(d, a0 + off, b0 + off, idesc_qk, ks > 0 ? 1u : 0u);
            }
          }
        }
      };
      // keys [64*hf, 64*hf+64) of the block -> S columns [64*hf, 64*hf+64)
      auto qk_half = [&](int t, int slot, int hf) {
        const uint64_t a0 = dq0 + static_cast<uint64_t>((t * S::kTileBytes) >> 4);
        const uint64_t b0 = dk0 + static_cast<uint64_t>(((slot % kNStage) * S::kTileBytes + hf * (64 * 128)) >> 4);
        const uint32_t d = tm + t * 128 + hf * 64;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t off = ((ks >> 2) * S::kBoxBytes + (ks & 3) * 32) >> 4;
            umma_ss(d, a0 + off, b0 + off, idesc_qk64, ks > 0 ? 1u : 0u);
          }
        }
      };
      // O_t += P_t[:, 64*hf .. 64*hf+64) V[64*hf .. 64*hf+64, :]
      auto pv_half = [&](int t, int slot, int hf, bool acc) {
        const uint64_t b0 = dv0 + static_cast<uint64_t>(((slot % kNStage) * S::kTileBytes) >> 4);
        const uint32_t d = tm + 256 + (SINGLE ? 0 : t * D);
        // P aliases S columns [0, 64); QUAD: keys 64.. sit in the first 32 columns of the second half of S
        const uint32_t a = tm + t * 128 + ((QUAD && hf) ? 32 : 0);
        if (elect_one()) {
#pragma unroll
          for (int j = 4 * hf; j < 4 * hf + 4; ++j) {
            umma_ts(d, a + j * 8, b0 + ((j * 2048) >> 4), idesc_pv, (acc || j > 0) ? 1u : 0u);
            if constexpr (SPLIT)  // the low term of P (S columns 64..) against the same [Vh | Vl] rows
              umma_ts(d, a + 64 + j * 8, b0 + ((j * 2048) >> 4), idesc_pv, 1u);
          }
        }
      };
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) tc_commit(bar);
      };
      if constexpr (SINGLE) {
        // One q tile, TWO S buffers (TMEM columns 0..127 / 128..255): QK(i+1) is issued before PV(i), so the
        // tensor pipe computes the next block's logits while the softmax warpgroup works on the current one --
        // the overlap the two-tile kernels get from their second tile.  K always sits in ring slot 0, V in
        // slot 1.  In-order execution keeps QK(i+1) (which overwrites P(i-1)) behind PV(i-1).  o_full[1]
        // completes a phase per PV: the softmax warpgroup waits on it before it rescales O.
        mbar_wait(&q_full[0], 0);
        slot_wait(0);
        tc_fence_after();
        qk(0, 0);
        commit(&s_full[0]);
        commit(&kv_empty[0]);
        FWD_STAMP(0, 10);
        for (int i = 0; i < nblk; ++i) {
          const int bsel = i & 1, vslot = 2 * i + 1, knext = 2 * i + 2;
          if (i + 1 < nblk) {
            slot_wait(knext);
            tc_fence_after();
            FWD_STAMP(i, 11);
            qk(bsel ^ 1, knext);
            commit(&s_full[bsel ^ 1]);
            commit(&kv_empty[knext % kNStage]);
          }
          mbar_wait(&p_half[2 * bsel], (i >> 1) & 1);
          FWD_STAMP(i, 12);
          slot_wait(vslot);
          tc_fence_after();
          FWD_STAMP(i, 13);
          pv_half(bsel, vslot, 0, i > 0);
          mbar_wait(&p_half[2 * bsel + 1], (i >> 1) & 1);
          tc_fence_after();
          pv_half(bsel, vslot, 1, true);
          commit(&kv_empty[vslot % kNStage]);
          commit(&o_full[1]);
          if (i + 1 == nblk) commit(&o_full[0]);
          FWD_STAMP(i, 14);
        }
      } else {
      mbar_wait(&q_full[0], 0);
      slot_wait(0);
      tc_fence_after();
      qk(0, 0);
      commit(&s_full[0]);
      if (act1) {
        mbar_wait(&q_full[1], 0);
        tc_fence_after();
        qk(1, 0);
        commit(&s_full[1]);
      }
      commit(&kv_empty[0]);
      // One issuer walks both tiles in turn.  That serialisation is what keeps the two tiles half a
      // step apart (one tile's softmax under the other's MMAs): with one issuer warp per tile
      // (tried) both chains fall into lockstep -- both softmaxes contend for the same issue slots
      // and then both MMA groups queue on the pipe -- and the step grows from 3 200 to 4 490 clk.
      // Per step and tile the tensor pipe sees PV_lo(i), PV_hi(i) (as each half of P lands) and
      // QK(i+1) (its columns alias P, so it follows PV in the in-order pipe).  With kSplitQK the
      // upper half of S(i+1) is issued as soon as the softmax warpgroup holds S(i) in registers.
      for (int i = 0; i < nblk; ++i) {
        const int vslot = 2 * i + 1, knext = 2 * i + 2;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int nbt = t ? nb1 : nb0;
          if (i >= nbt) continue;
          const bool has_next = i + 1 < nbt;
          if (kSplitQK && has_next) {
            mbar_wait(&s_read[t], i & 1);
            slot_wait(knext);
            tc_fence_after();
            qk_half(t, knext, 1);
          }
          mbar_wait(&p_half[2 * t], i & 1);
          if (t == 0 && !FWD_TRACE_WARPS) FWD_STAMP(i, 14);
          slot_wait(vslot);
          tc_fence_after();
          if (!FWD_TRACE_WARPS) FWD_STAMP(i, 10 + 2 * t);
          pv_half(t, vslot, 0, i > 0);
          if (kSplitPV) {
            mbar_wait(&p_half[2 * t + 1], i & 1);
            tc_fence_after();
            if (t == 0 && !FWD_TRACE_WARPS) FWD_STAMP(i, 15);
          }
          pv_half(t, vslot, 1, true);
          if (has_next) {
            if (kSplitQK) {
              qk_half(t, knext, 0);
            } else {
              slot_wait(knext);
              tc_fence_after();
              qk(t, knext);
            }
            commit(&s_full[t]);
          } else {
            commit(&o_full[t]);
          }
          if (!FWD_TRACE_WARPS) FWD_STAMP(i, 11 + 2 * t);
        }
        commit(&kv_empty[vslot % kNStage]);
        if (i + 1 < nblk) commit(&kv_empty[knext % kNStage]);
      }
      }
    }
  } else if constexpr (QUAD) {
    // ====================== softmax warps, two per 32 rows of a tile =======================
    setmaxnreg_inc<104>();   // 640 x 96 at launch; the producer warpgroup's 128 x 32 cover 512 x 8
    const int sw = warp - kFirstSoftmaxWarp;          // 0..15
    const int t = sw >> 3, hf = (sw >> 2) & 1;         // tile, half of the key columns
    const int nbt = t ? nb1 : nb0;
    const int wq = warp & 3;                           // TMEM lane quarter this warp may touch
    const int row = wq * 32 + lane;
    const int q_row = q0 + t * 128 + row;
    const uint32_t pair_bar = 3 + t * 4 + wq;          // named barrier of the two warps that share these rows
    float* xmax = reinterpret_cast<float*>(smem + ((S::kTotal + 15) & ~15)) + t * 512;   // [parity][half][row]
    float* xsum = reinterpret_cast<float*>(smem + ((S::kTotal + 15) & ~15)) + 1024 + t * 256;  // [half][row]
    if (nbt > 0) {
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tS = tmem_base + lane_off + t * 128 + hf * 64;       // my 64 logits; my P = its first 32 columns
      const uint32_t tO = tmem_base + lane_off + 256 + t * D + hf * (D / 2);  // my half of O's columns
      const float sl2 = p.scale_log2;
      const uint64_t sl2x2 = pack_f2(sl2, sl2);
      float m_used = -1e30f;   // reference max, scaled log2 units; identical in both threads of a row
      float l = 0.f;           // partial row sum (my 64 columns of every block)
      auto mask_bytes = [&](int blk) -> uint32_t {
        uint32_t r = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kk = blk * 128 + 4 * lane + j;
          if (kk < KL && kmask[kk]) r |= 1u << (8 * j);
        }
        return r;
      };
      uint32_t mb_next = kmask ? mask_bytes(0) : 0u;
      for (int i = 0; i < nbt; ++i) {
        const uint32_t mb = mb_next;
        if (kmask && i + 1 < nbt) mb_next = mask_bytes(i + 1);
        mbar_wait(&s_full[t], i & 1);
        tc_fence_after();
        if ((sw & 7) == 0) FWD_STAMP(i, 5 * t + 0);
        uint32_t sr[2][32];
        tmem_ld_x32(tS, sr[0]);
        tmem_ld_x32(tS + 32, sr[1]);
        tmem_ld_wait();
        if ((sw & 7) == 0) FWD_STAMP(i, 5 * t + 1);
        const int k0 = i * 128;
        uint32_t kw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (kmask) {
#pragma unroll
          for (int j = 0; j < 4; ++j) kw[j] = __ballot_sync(0xffffffffu, (mb >> (8 * j)) & 1u);
        }
        const bool need_mask = (k0 + 128 > KL) || (p.causal && (k0 + 127 > q0 + t * 128)) ||
                               ((kw[0] & kw[1] & kw[2] & kw[3]) != 0xffffffffu);
        if (need_mask) {
          const int lim = p.causal ? min(KL - 1, q_row) : (KL - 1);  // last visible key
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int kk = hf * 64 + c * 32 + j;
              if (k0 + kk > lim || !((kw[kk & 3] >> (kk >> 2)) & 1u)) sr[c][j] = 0xff800000u;  // -inf
            }
        }
        // partial row max of my 64 columns -> shared memory for the other half's thread
        float mx8[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            mx8[j & 3] = fmax3(mx8[j & 3], __uint_as_float(sr[c][2 * j]), __uint_as_float(sr[c][2 * j + 1]));
        const float mx_mine = fmax3(mx8[0], mx8[1], fmaxf(mx8[2], mx8[3]));
        xmax[(i & 1) * 256 + hf * 128 + row] = mx_mine;
        // exp2(S * scale * log2e - m_used) of 16 columns (sub-chunk sc of my 64), packed to 16 bits, summed
        auto exp_sub = [&](int sc, uint32_t (&pr)[8], uint64_t& sum2) {
          const uint64_t negm2 = pack_f2(-m_used, -m_used);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = (sc & 1) * 16 + 2 * j;
            const uint64_t x2 = ffma2(pack_f2(__uint_as_float(sr[sc >> 1][col]), __uint_as_float(sr[sc >> 1][col + 1])),
                                      sl2x2, negm2);
            float x0, x1;
            unpack_f2(x2, x0, x1);
            float p0, p1;
            if (kQuadPolyEvery > 0 && (j % (kQuadPolyEvery > 0 ? kQuadPolyEvery : 1)) == kQuadPolyEvery - 1) {
              exp2_poly2(x0, x1, p0, p1);
            } else {
              p0 = fast_exp2(x0);
              p1 = fast_exp2(x1);
            }
            sum2 = fadd2(sum2, pack_f2(p0, p1));
            pr[j] = pack2<T>(p0, p1);
          }
        };
        // speculation (see the two-warpgroup kernel): the first sub-chunk is exponentiated against the running
        // reference max before the barrier, i.e. while the other half's partial max is on its way
        uint32_t pr0[8];
        uint64_t sum0 = pack_f2(0.f, 0.f);
        exp_sub(0, pr0, sum0);
        named_bar_sync(pair_bar, 64);
        const float mx = fmaxf(mx_mine, xmax[(i & 1) * 256 + (hf ^ 1) * 128 + row]);
        const float mx_s = mx * sl2;
        const bool grow = mx_s > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, grow)) {   // same rows, same inputs: the partner warp takes the same branch
          const float m_new = grow ? mx_s : m_used;
          const float alpha = fast_exp2(m_used - m_new);
          m_used = m_new;
          l *= alpha;
          if (i > 0) {
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {   // 16 columns at a time: the 64 logits stay in registers
              uint32_t orow[16];
              tmem_ld_x16(tO + c * 16, orow);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) orow[j] = __float_as_uint(__uint_as_float(orow[j]) * alpha);
              tmem_st_x16(tO + c * 16, orow);
            }
            // PV(i) adds into ALL of O's columns as soon as either half of P is announced: both halves of O
            // must have been rescaled by then
            tmem_st_wait();
            tc_fence_before();
            named_bar_sync(pair_bar, 64);
          }
          sum0 = pack_f2(0.f, 0.f);
          exp_sub(0, pr0, sum0);
        }
        if ((sw & 7) == 0) FWD_STAMP(i, 5 * t + 2);
        tmem_st_x8(tS, pr0);
        uint64_t sum2 = sum0;
#pragma unroll
        for (int sc = 1; sc < 4; ++sc) {
          uint32_t pr[8];
          exp_sub(sc, pr, sum2);
          tmem_st_x8(tS + sc * 8, pr);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_half[2 * t + hf]);
        if ((sw & 3) == 0) FWD_STAMP(i, 5 * t + 3 + hf);
        float s0, s1;
        unpack_f2(sum2, s0, s1);
        l += s0 + s1;
      }

      // ---- epilogue: my 64 columns of O / l -> 16-bit -> box `hf` of the swizzled Q_t buffer -> TMA store ----
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      xsum[hf * 128 + row] = l;
      named_bar_sync(pair_bar, 64);
      l += xsum[(hf ^ 1) * 128 + row];
      const float inv_l = l > 0.f ? 1.f / l : 0.f;
      uint8_t* stage = sQ + t * S::kTileBytes + hf * S::kBoxBytes;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
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
          const int cin = c * 4 + u;  // 16-byte chunk within the 128-byte box row
          *reinterpret_cast<uint4*>(stage + row * 128 + ((cin ^ (row & 7)) << 4)) = v;
        }
      }
      if (hf == 0 && q_row < QL)
        p.lse[static_cast<int64_t>(bh_q) * QL + q_row] = l > 0.f ? (m_used + fast_log2(l)) * kLn2 : -INFINITY;
      fence_proxy_async_smem();
      named_bar_sync(1 + t, 256);
      if ((sw & 7) == 0 && lane == 0) {
#pragma unroll
        for (int bx = 0; bx < S::kNBox; ++bx)
          tma_store_3d(&tm_o, sQ + t * S::kTileBytes + bx * S::kBoxBytes, bx * 64, q_off + q0 + t * 128, bh_q);
        bulk_commit();
        bulk_wait_read<0>();
      }
    } else if (hf == 0 && (t == 0 || act1)) {
      // no visible keys at all (KL == 0): the output rows are 0 and lse = -inf
      if (q_row < QL) {
        const int64_t ri = static_cast<int64_t>(bh_q) * QL + q_row;
        uint4* orow = reinterpret_cast<uint4*>(static_cast<char*>(p.o_ptr) + ri * p.o_row_bytes);
        for (int c = 0; c < p.o_row_bytes / 16; ++c) orow[c] = make_uint4(0u, 0u, 0u, 0u);
        p.lse[ri] = -INFINITY;
      }
    }
  } else {
    // ================================ softmax warpgroups ===============================
    setmaxnreg_inc<208>();
    const int t = (warp - 4) >> 2;
    const int nbt = t ? nb1 : nb0;
    if (nbt > 0) {
      const int wq = warp & 3;
      const int row = wq * 32 + lane;
      const int q_row = q0 + t * 128 + row;
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tS0 = tmem_base + lane_off + t * 128;
      const uint32_t tO = tmem_base + lane_off + 256 + t * D;
      // with a bias the logits are moved to log2 units as they are folded, so the rest runs unscaled
      // Float32 (SPLIT): q', k' are the caller's q, k times exact powers of two (scale block); the logit scale
      // takes 2^(e_q + e_k) back, the epilogue 2^e_v
      const float scale_log2 = SPLIT ? p.scale_log2 * __ldg(p.f32_mult + F32Mult::kLogits) : p.scale_log2;
      const float sl2 = BIAS ? 1.f : scale_log2;
      float m_used = -1e30f;  // reference max in scaled log2 units (finite: see the speculation note)
      float l = 0.f;

      const uint64_t sl2x2 = pack_f2(sl2, sl2);
      // key padding: lane holds mask bytes 4*lane .. 4*lane+3 of a block; fetched one block ahead
      auto mask_bytes = [&](int blk) -> uint32_t {
        uint32_t r = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kk = blk * 128 + 4 * lane + j;
          if (kk < KL && kmask[kk]) r |= 1u << (8 * j);
        }
        return r;
      };
      uint32_t mb_next = kmask ? mask_bytes(0) : 0u;
      for (int i = 0; i < nbt; ++i) {
        const uint32_t mb = mb_next;
        if (kmask && i + 1 < nbt) mb_next = mask_bytes(i + 1);
        // SINGLE: the logits alternate between two TMEM buffers (block i -> buffer i & 1, a phase every other block)
        const int bsel = SINGLE ? (i & 1) : t;
        const uint32_t bph = SINGLE ? ((i >> 1) & 1) : (i & 1);
        const uint32_t tS = SINGLE ? tS0 + bsel * 128 : tS0;
        mbar_wait(&s_full[bsel], bph);
        tc_fence_after();
        if (wq == 0 && !FWD_TRACE_WARPS) FWD_STAMP(i, 5 * t + 0);
        uint32_t sr[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, sr[c]);
        tmem_ld_wait();
        if (FWD_TRACE_WARPS ? t == 0 : wq == 0) FWD_STAMP(i, FWD_TRACE_WARPS ? 4 * wq : 5 * t + 1);
        if (kSplitQK) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_read[t]);
        }
        if constexpr (BIAS) {
          // S <- S * scale * log2e + pair * log2e, one 128-byte box row (this thread's query) at a time
          const float sraw = scale_log2;
#pragma unroll
          for (int c = 0; c < kBiasChunks; ++c) {
            const int n = i * kBiasChunks + c;
            const int u = 2 * t + (n & 1);
            mbar_wait(&bias_full[u], (n >> 1) & 1);
            const uint8_t* brow = sBias + u * 16384 + row * 128;
#pragma unroll
            for (int u16 = 0; u16 < 8; ++u16) {
              const uint4 bv = *reinterpret_cast<const uint4*>(brow + ((u16 ^ (row & 7)) << 4));
              if constexpr (sizeof(BT) == 4) {
                const int col = c * 32 + u16 * 4;
                const uint32_t w[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  uint32_t& sv = sr[(col + e) >> 5][(col + e) & 31];
                  sv = __float_as_uint(fmaf(__uint_as_float(sv), sraw, __uint_as_float(w[e]) * kLog2e));
                }
              } else {
                const int col = c * 64 + u16 * 8;
                const uint32_t w[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  uint32_t& s0 = sr[(col + 2 * e) >> 5][(col + 2 * e) & 31];
                  uint32_t& s1 = sr[(col + 2 * e + 1) >> 5][(col + 2 * e + 1) & 31];
                  s0 = __float_as_uint(fmaf(__uint_as_float(s0), sraw, unpack_lo<T>(w[e]) * kLog2e));
                  s1 = __float_as_uint(fmaf(__uint_as_float(s1), sraw, unpack_hi<T>(w[e]) * kLog2e));
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bias_empty[u]);
          }
        }

        const int k0 = i * 128;
        // bit l of kw[j] = mask of key 4*l + j of this block (all ones without a padding mask)
        uint32_t kw[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (kmask) {
#pragma unroll
          for (int j = 0; j < 4; ++j) kw[j] = __ballot_sync(0xffffffffu, (mb >> (8 * j)) & 1u);
        }
        const bool need_mask = (k0 + 128 > KL) || (p.causal && (k0 + 127 > q0 + t * 128)) ||
                               ((kw[0] & kw[1] & kw[2] & kw[3]) != 0xffffffffu);
        if (need_mask) {
          const int lim = p.causal ? min(KL - 1, q_row) : (KL - 1);  // last visible key
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int kk = c * 32 + j;
              if (k0 + kk > lim || !((kw[kk & 3] >> (kk >> 2)) & 1u)) sr[c][j] = 0xff800000u;  // -inf
            }
        }
        // exp2(S*scale*log2e - m_used) of one 32-key chunk; every kPolyEvery-th pair runs on the
        // FMA pipe (exp2_poly2) instead of the MUFU
        auto exp_chunk = [&](int c, float (&pf)[32]) {
          const uint64_t negm2 = pack_f2(-m_used, -m_used);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint64_t x2 = ffma2(pack_f2(__uint_as_float(sr[c][2 * j]), __uint_as_float(sr[c][2 * j + 1])),
                                      sl2x2, negm2);
            float x0, x1;
            unpack_f2(x2, x0, x1);
            if (kPolyEvery > 0 && (j % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1) {
              exp2_poly2(x0, x1, pf[2 * j], pf[2 * j + 1]);
            } else {
#ifdef NNOP_FWD_NO_EXP   // timing experiment only (wrong results): what do the MUFU exponentials cost?
              pf[2 * j] = x0;
              pf[2 * j + 1] = x1;
#else
              pf[2 * j] = fast_exp2(x0);
              pf[2 * j + 1] = fast_exp2(x1);
#endif
            }
          }
        };
        float pf[32];
        // Speculation: start exponentiating chunk 0 against the running reference max while the
        // row max of this block is still being reduced (ALU pipe, off the critical path).  Any
        // finite reference is exact as long as P, l and O share it; only if the block raises the
        // max by more than 2^kRescaleThreshold is the reference moved (O, l rescaled) and chunk 0
        // redone.  m_used starts at -1e30 so the first block always takes that path.
        exp_chunk(0, pf);
        float mx8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) mx8[u] = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            mx8[j & 7] = fmax3(mx8[j & 7], __uint_as_float(sr[c][2 * j]), __uint_as_float(sr[c][2 * j + 1]));
        const float mx = fmax3(fmax3(mx8[0], mx8[1], mx8[2]), fmax3(mx8[3], mx8[4], mx8[5]),
                               fmaxf(mx8[6], mx8[7]));
        const float mx_s = mx * sl2;
        const bool grow = mx_s > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mx_s : m_used;
          const float alpha = fast_exp2(m_used - m_new);
          m_used = m_new;
          l *= alpha;
          if (i > 0) {
            if constexpr (SINGLE) {   // S(i) was issued ahead of PV(i-1): O is only stable once that PV has completed
              mbar_wait(&o_full[1], (i - 1) & 1);
              tc_fence_after();
            }
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
          exp_chunk(0, pf);
        }
        if (FWD_TRACE_WARPS ? t == 0 : wq == 0) FWD_STAMP(i, FWD_TRACE_WARPS ? 4 * wq + 1 : 5 * t + 2);
        uint64_t sum2 = pack_f2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c > 0) exp_chunk(c, pf);
          uint32_t pr[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            sum2 = fadd2(sum2, pack_f2(pf[2 * j], pf[2 * j + 1]));
            pr[j] = pack2<T>(pf[2 * j], pf[2 * j + 1]);
          }
          tmem_st_x16(tS + c * 16, pr);
          if constexpr (SPLIT) {  // low term: p - bf16(p), also 16-bit
            uint32_t plo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              plo[j] = pack2<T>(pf[2 * j] - unpack_lo<T>(pr[j]), pf[2 * j + 1] - unpack_hi<T>(pr[j]));
            tmem_st_x16(tS + 64 + c * 16, plo);
          }
          if (kSplitPV ? (c & 1) : (c == 3)) {  // (half of) the keys are in TMEM: release the tensor pipe
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_half[2 * bsel + (kSplitPV ? (c >> 1) : 0)]);
            if (FWD_TRACE_WARPS ? t == 0 : wq == 0) FWD_STAMP(i, FWD_TRACE_WARPS ? 4 * wq + 2 + (c >> 1) : 5 * t + 3 + (c >> 1));
          }
        }
        float s0, s1;
        unpack_f2(sum2, s0, s1);
        l += s0 + s1;
      }

      // ---- epilogue: O / l -> 16-bit -> swizzled smem (the Q_t buffer) -> TMA store -----
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      float inv_l = l > 0.f ? 1.f / l : 0.f;
      uint8_t* stage = sQ + t * S::kTileBytes;
      if constexpr (SPLIT) {
        inv_l *= __ldg(p.f32_mult + F32Mult::kO);   // O = 2^e_v * P V'
        // O'[:, 0:D/2] = P Vh, O'[:, D/2:D] = P Vl: add them, normalise, stage as fp32 (boxes of
        // 32 floats x 128 rows, same 128-byte swizzle) for the fp32 TMA store
#pragma unroll
        for (int c = 0; c < D / 64; ++c) {
          uint32_t ohi[32], olo[32];
          tmem_ld_x32(tO + c * 32, ohi);
          tmem_ld_x32(tO + D / 2 + c * 32, olo);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 v;
            v.x = (__uint_as_float(ohi[4 * u + 0]) + __uint_as_float(olo[4 * u + 0])) * inv_l;
            v.y = (__uint_as_float(ohi[4 * u + 1]) + __uint_as_float(olo[4 * u + 1])) * inv_l;
            v.z = (__uint_as_float(ohi[4 * u + 2]) + __uint_as_float(olo[4 * u + 2])) * inv_l;
            v.w = (__uint_as_float(ohi[4 * u + 3]) + __uint_as_float(olo[4 * u + 3])) * inv_l;
            *reinterpret_cast<float4*>(stage + c * S::kBoxBytes + row * 128 + ((u ^ (row & 7)) << 4)) = v;
          }
        }
      } else {
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
      }
      if (q_row < QL) {
        const int64_t li = packed ? static_cast<int64_t>(h) * p.total_q + q_off + q_row
                                  : static_cast<int64_t>(bh_q) * QL + q_row;
        p.lse[li] = l > 0.f ? (m_used + fast_log2(l)) * kLn2 : -INFINITY;
      }
      const int rows_left = QL - (q0 + t * 128);  // > 0 here
      if (packed && rows_left < 128) {
        // last, partial tile of a packed sequence: rows past its end belong to the next sequence,
        // so copy the valid rows out of the staging buffer with coalesced 16-byte stores
        named_bar_sync(1 + t, 128);
        constexpr int kCPR = D / 8;  // 16-byte chunks per row
        T* obase = static_cast<T*>(p.o_ptr) +
                   (static_cast<int64_t>(h) * p.total_q + q_off + q0 + t * 128) * D;
        const int tid = wq * 32 + lane;
        for (int idx = tid; idx < rows_left * kCPR; idx += 128) {
          const int r = idx / kCPR, chunk = idx % kCPR;
          const int bx = chunk >> 3, cin = chunk & 7;
          const uint4 v = *reinterpret_cast<const uint4*>(stage + bx * S::kBoxBytes + r * 128 +
                                                          ((cin ^ (r & 7)) << 4));
          *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(r) * D + chunk * 8) = v;
        }
      } else {
        fence_proxy_async_smem();
        named_bar_sync(1 + t, 128);
        if (wq == 0 && lane == 0) {
#pragma unroll
          for (int bx = 0; bx < S::kNBox; ++bx)
            tma_store_3d(&tm_o, stage + bx * S::kBoxBytes, bx * (SPLIT ? 32 : 64), q_off + q0 + t * 128, bh_q);
          bulk_commit();
          bulk_wait_read<0>();
        }
      }
    } else if (t == 0 || act1) {
      // no visible keys at all (KL == 0): the output rows are 0 and lse = -inf
      const int q_row = q0 + t * 128 + (warp & 3) * 32 + lane;
      if (q_row < QL) {
        const int64_t ri = packed ? static_cast<int64_t>(h) * p.total_q + q_off + q_row
                                  : static_cast<int64_t>(bh_q) * QL + q_row;
        uint4* orow = reinterpret_cast<uint4*>(static_cast<char*>(p.o_ptr) + ri * p.o_row_bytes);
        for (int c = 0; c < p.o_row_bytes / 16; ++c) orow[c] = make_uint4(0u, 0u, 0u, 0u);
        p.lse[ri] = -INFINITY;
      }
    }
  }

  // ---- teardown -------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// =========================================================================================
// Persistent forward (dense layout, bf16 / f16, no key padding mask, no bias): one CTA per SM walks a
// dynamic queue of (256-row q tile, q head, batch) work tiles, heaviest q tile of a head first and
// the q tiles of one head next to each other (they share K / V through L2).  The one-tile kernel
// above pays ~7 K clk per CTA outside its steady state (launch gap, barrier / TMEM setup, the Q and
// first K load, the epilogue) and with one CTA per SM nothing overlaps it; here
//   * the K / V ring, the barrier phases and TMEM run on across tiles: the producer loads the next
//     tile's Q_t as soon as the last QK of the current tile has read Q_t, and its first K / V blocks
//     behind the current tile's last ones;
//   * the next tile's first QK is issued right behind the last PV; only its first PV waits for the
//     softmax warpgroup to have pulled O out of TMEM (o_empty);
//   * each softmax warpgroup stores its O tile through a private 16 KB staging buffer (one 64-column
//     box at a time), so Q never waits for a store;
//   * warp 3 claims tiles with atomicAdd on a counter in the caller's workspace (see launch_fwd_persist) and
//     publishes them through a two-slot ring, late, as in the persistent backward.
// The softmax code is the one-tile kernel's, so O and lse are bit-identical (test).
// =========================================================================================
template <int D>
struct FwdPersistSmem {
  static constexpr int kNStage = 4;
  static constexpr int kTileBytes = 128 * D * 2;
  static constexpr int kBoxBytes = 128 * 64 * 2;
  static constexpr int kNBox = D / 64;
  static constexpr int kQOff = 0;
  static constexpr int kKVOff = 2 * kTileBytes;
  static constexpr int kStageOff = kKVOff + kNStage * kTileBytes;  // 2 x 16 KB: O staging per warpgroup
  static constexpr int kBarOff = kStageOff + 2 * kBoxBytes;
  enum : int {
    kQFull = 0,      // [2]
    kQEmpty = 2,     // [2]
    kKVFull = 4,     // [4]
    kKVEmpty = 8,    // [4]
    kSFull = 12,     // [2]
    kPHalf = 14,     // [2][2]
    kOFull = 18,     // [2]
    kOEmpty = 20,    // [2]
    kTileFull = 22,  // [2]
    kTileEmpty = 24, // [2]
    kSchedGo = 26,
    kNumBars = 27
  };
  static constexpr int kTileRing = kBarOff + kNumBars * 8;   // [2][4] ints
  static constexpr int kTmemSlot = kTileRing + 32;
  static constexpr int kSeqChunks = 128;                     // packed mode: coarse prefix of q tiles
  static constexpr int kSeqPre = kTmemSlot + 16;             // [kSeqChunks + 1] ints
  static constexpr int kTotal = kSeqPre + (kSeqChunks + 1) * 4 + 12;
  static constexpr int kDynBytes = kTotal + 1024;
};

template <typename T, int D>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_sm100_persist_kernel(const __grid_constant__ CUtensorMap tm_q,
                              const __grid_constant__ CUtensorMap tm_k,
                              const __grid_constant__ CUtensorMap tm_v,
                              const __grid_constant__ CUtensorMap tm_o, const FwdParams p,
                              int* __restrict__ tile_counter, const int n_tiles, const int nqt) {
  using S = FwdPersistSmem<D>;
  constexpr int kNStage = S::kNStage;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + S::kQOff;
  uint8_t* sKV = smem + S::kKVOff;
  uint8_t* sStage = smem + S::kStageOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  volatile int* s_tile = reinterpret_cast<volatile int*>(smem + S::kTileRing);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S::kTmemSlot);

  int* s_pre = reinterpret_cast<int*>(smem + S::kSeqPre);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const bool packed = p.cu_q != nullptr;
  // packed batches: q tiles per sequence, summed per chunk of sequences (exclusive prefix in s_pre)
  const int seq_per = packed ? (p.nseq + S::kSeqChunks - 1) / S::kSeqChunks : 0;
  auto tiles_of = [&](int z) { return (p.cu_q[z + 1] - p.cu_q[z] + 255) >> 8; };
  if (packed) {
    if (threadIdx.x < S::kSeqChunks) {
      const int z0 = min(p.nseq, static_cast<int>(threadIdx.x) * seq_per), z1 = min(p.nseq, z0 + seq_per);
      int cnt = 0;
      for (int z = z0; z < z1; ++z) cnt += tiles_of(z);
      s_pre[threadIdx.x + 1] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      s_pre[0] = 0;
      for (int c = 0; c < S::kSeqChunks; ++c) s_pre[c + 1] += s_pre[c];
    }
    // (made visible to everyone by the __syncthreads after the barrier setup below)
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
    for (int i = 0; i < S::kNumBars; ++i) {
      uint32_t cnt = 1;
      if (i >= S::kPHalf && i < S::kPHalf + 4) cnt = 4;     // one arrival per softmax warp
      if (i >= S::kOEmpty && i < S::kOEmpty + 2) cnt = 4;
      if (i >= S::kTileEmpty && i < S::kTileEmpty + 2) cnt = 10;  // producer + issuer + 8 softmax warps
      mbar_init(bars + i, cnt);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work tile geometry, derived from a ring record by every role the same way
  struct Tile {
    int bh_q, bh_kv, q0, act1, nb0, nb1, nblk, QL, KL, q_off, k_off;
  };
  // zb = batch element (dense) or sequence (packed)
  auto make_tile = [&](int zb, int h, int qt) -> Tile {
    Tile t;
    t.q0 = qt * 256;
    if (packed) {
      t.q_off = p.cu_q[zb];
      t.QL = p.cu_q[zb + 1] - t.q_off;
      t.k_off = p.cu_k[zb];
      t.KL = p.cu_k[zb + 1] - t.k_off;
      t.bh_q = h;
      t.bh_kv = h / (p.QH / p.KH);
    } else {
      t.q_off = 0; t.k_off = 0; t.QL = p.QL; t.KL = p.KL;
      t.bh_q = zb * p.QH + h;
      t.bh_kv = zb * p.KH + h / (p.QH / p.KH);
    }
    t.act1 = t.q0 + 128 < t.QL;
    t.nb0 = ((p.causal ? min(t.KL, t.q0 + 128) : t.KL) + 127) >> 7;
    t.nb1 = t.act1 ? (((p.causal ? min(t.KL, t.q0 + 256) : t.KL) + 127) >> 7) : 0;
    t.nblk = t.act1 ? t.nb1 : t.nb0;
    return t;
  };
  auto next_tile = [&](int n, Tile& t) -> bool {
    const int slot = n & 1;
    mbar_wait(bars + S::kTileFull + slot, (n >> 1) & 1);
    const int b = s_tile[4 * slot], h = s_tile[4 * slot + 1], qt = s_tile[4 * slot + 2];
    const int flag = s_tile[4 * slot + 3];
    __syncwarp();
    if (lane == 0) mbar_arrive(bars + S::kTileEmpty + slot);
    if (flag < 0) return false;
    t = make_tile(b, h, qt);
    return true;
  };

  if (warp < 4) {
    setmaxnreg_dec<88>();
    if (warp == 3) {
      // ================================ tile scheduler ===============================
      // Dense: tile t = (batch * QH + head) * nqt + r.  Packed: sequences in order, within a sequence
      // head-major with the q tiles of a head adjacent; the sequence is found in the coarse prefix, then
      // by a walk over its chunk.  The whole warp runs this (uniformly); lane 0 claims and publishes.
      const int total = packed ? s_pre[S::kSeqChunks] * p.QH : n_tiles;
      for (int n = 0;; ++n) {
        const int slot = n & 1;
        if (n > 0) mbar_wait(bars + S::kSchedGo, (n - 1) & 1);   // claim late (see the backward)
        mbar_wait(bars + S::kTileEmpty + slot, ((n >> 1) & 1) ^ 1);
        int zb = 0, h = 0, qt = 0, flag = -1;
        for (;;) {
          int t = 0;
          if (lane == 0) t = atomicAdd(tile_counter, 1);
          t = __shfl_sync(0xffffffffu, t, 0);
          if (t >= total) break;
          int nq_z = nqt, r;
          if (packed) {
            // t = QH * (tiles before sequence z) + h * nq_z + r  ->  find z with QH * pre_z <= t
            int c = 0;
            while (c + 1 < S::kSeqChunks && s_pre[c + 1] * p.QH <= t) ++c;
            int z = c * seq_per, run = s_pre[c];
            nq_z = tiles_of(z);
            while ((run + nq_z) * p.QH <= t) {
              run += nq_z;
              nq_z = tiles_of(++z);
            }
            const int tl = t - run * p.QH;
            zb = z;
            h = tl / nq_z;
            r = tl - h * nq_z;
          } else {
            const int bh = t / nqt;
            r = t - bh * nqt;
            zb = bh / p.QH;
            h = bh - zb * p.QH;
          }
          qt = p.causal ? nq_z - 1 - r : r;   // heaviest q tile of a head first
          const Tile ti = make_tile(zb, h, qt);
          if (ti.nb0 > 0) { flag = 1; break; }
          // a packed sequence without keys: its output rows are 0 and lse = -inf (written here with plain
          // stores; the tile is not published)
          {
            const int rows = min(256, ti.QL - ti.q0);
            constexpr int kCPR = D / 8;
            const int64_t row0 = static_cast<int64_t>(h) * p.total_q + ti.q_off + ti.q0;
            for (int idx = lane; idx < rows * kCPR; idx += 32)
              reinterpret_cast<uint4*>(static_cast<T*>(p.o_ptr) + row0 * D)[idx] = make_uint4(0u, 0u, 0u, 0u);
            for (int idx = lane; idx < rows; idx += 32) p.lse[row0 + idx] = -INFINITY;
          }
        }
        if (lane == 0) {
          s_tile[4 * slot] = zb; s_tile[4 * slot + 1] = h; s_tile[4 * slot + 2] = qt; s_tile[4 * slot + 3] = flag;
          mbar_arrive(bars + S::kTileFull + slot);
        }
        if (flag < 0) break;
      }
    } else if (warp == 0) {
      // ================================ TMA producer =================================
      if (lane == 0) {
        int n = 0;        // K / V loads issued so far (ring position)
        int cq[2] = {0, 0};  // uses of Q buffer t so far
        for (int tl = 0;; ++tl) {
          const int slot = tl & 1;
          mbar_wait(bars + S::kTileFull + slot, (tl >> 1) & 1);
          const int tb = s_tile[4 * slot], th = s_tile[4 * slot + 1], tqt = s_tile[4 * slot + 2];
          const int flag = s_tile[4 * slot + 3];
          mbar_arrive(bars + S::kTileEmpty + slot);
          if (flag < 0) break;
          const Tile ti = make_tile(tb, th, tqt);
          auto load_q = [&](int t) {
            mbar_wait(bars + S::kQEmpty + t, (cq[t] & 1) ^ 1);
            mbar_arrive_expect_tx(bars + S::kQFull + t, S::kTileBytes);
#pragma unroll
            for (int bx = 0; bx < S::kNBox; ++bx)
              tma_load_3d(sQ + t * S::kTileBytes + bx * S::kBoxBytes, &tm_q, bars + S::kQFull + t, bx * 64,
                          ti.q_off + ti.q0 + t * 128, ti.bh_q);
            ++cq[t];
          };
          auto load_kv = [&](const CUtensorMap* tm, int blk) {
            const int st = n % kNStage;
            const uint32_t ph = (n / kNStage) & 1;
            mbar_wait(bars + S::kKVEmpty + st, ph ^ 1);
            mbar_arrive_expect_tx(bars + S::kKVFull + st, S::kTileBytes);
#pragma unroll
            for (int bx = 0; bx < S::kNBox; ++bx)
              tma_load_3d(sKV + st * S::kTileBytes + bx * S::kBoxBytes, tm, bars + S::kKVFull + st, bx * 64,
                          ti.k_off + blk * 128, ti.bh_kv);
            ++n;
          };
          const int go_at = ti.nblk > 2 ? ti.nblk - 2 : 0;  // block whose loads trigger the next claim
          if (go_at == 0) mbar_arrive(bars + S::kSchedGo);
          load_q(0);
          load_kv(&tm_k, 0);
          if (ti.act1) load_q(1);
          load_kv(&tm_v, 0);
          for (int i = 1; i < ti.nblk; ++i) {
            if (i == go_at) mbar_arrive(bars + S::kSchedGo);
            load_kv(&tm_k, i);
            load_kv(&tm_v, i);
          }
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ===================================
      constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, is_bf16<T>::value, false, false);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, D, is_bf16<T>::value, false, true);
      const uint32_t tm = uniform_u32(tmem_base);
      const uint32_t q_base = uniform_u32(smem_u32(sQ));
      const uint32_t kv_base = uniform_u32(smem_u32(sKV));
      const uint64_t dq0 = make_smem_desc_sw128(q_base, 16, 1024);
      const uint64_t dk0 = make_smem_desc_sw128(kv_base, 16, 1024);
      const uint64_t dv0 = make_smem_desc_sw128(kv_base, S::kBoxBytes, 1024);
      auto slot_wait = [&](int slot) {
        mbar_wait(bars + S::kKVFull + slot % kNStage, (slot / kNStage) & 1);
      };
      auto qk = [&](int t, int slot) {
        const uint64_t a0 = dq0 + static_cast<uint64_t>((t * S::kTileBytes) >> 4);
        const uint64_t b0 = dk0 + static_cast<uint64_t>(((slot % kNStage) * S::kTileBytes) >> 4);
        const uint32_t d = tm + t * 128;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t off = ((ks >> 2) * S::kBoxBytes + (ks & 3) * 32) >> 4;
            umma_ss(d, a0 + off, b0 + off, idesc_qk, ks > 0 ? 1u : 0u);
          }
        }
      };
      auto pv_half = [&](int t, int slot, int hf, bool acc) {
        const uint64_t b0 = dv0 + static_cast<uint64_t>(((slot % kNStage) * S::kTileBytes) >> 4);
        const uint32_t d = tm + 256 + t * D;
        const uint32_t a = tm + t * 128;  // P aliases S columns [0, 64)
        if (elect_one()) {
#pragma unroll
          for (int j = 4 * hf; j < 4 * hf + 4; ++j)
            umma_ts(d, a + j * 8, b0 + ((j * 2048) >> 4), idesc_pv, (acc || j > 0) ? 1u : 0u);
        }
      };
      auto commit = [&](int bar) {
        if (elect_one()) tc_commit(bars + bar);
      };
      int nbase = 0;          // ring position of this tile's K_0
      int cq[2] = {0, 0};     // tiles processed on q slot t
      int cs[2] = {0, 0};     // key blocks processed on q slot t (phases of s_full / p_half)
      for (int tl = 0;; ++tl) {
        Tile ti;
        if (!next_tile(tl, ti)) break;
        const int nbt[2] = {ti.nb0, ti.nb1};
        mbar_wait(bars + S::kQFull + 0, cq[0] & 1);
        slot_wait(nbase);
        tc_fence_after();
        qk(0, nbase);
        commit(S::kSFull + 0);
        if (ti.nb0 == 1) commit(S::kQEmpty + 0);   // Q_0 is not read again in this tile
        if (ti.act1) {
          mbar_wait(bars + S::kQFull + 1, cq[1] & 1);
          tc_fence_after();
          qk(1, nbase);
          commit(S::kSFull + 1);
          if (ti.nb1 == 1) commit(S::kQEmpty + 1);
        }
        commit(S::kKVEmpty + nbase % kNStage);
        for (int i = 0; i < ti.nblk; ++i) {
          const int vslot = nbase + 2 * i + 1, knext = nbase + 2 * i + 2;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            if (i >= nbt[t]) continue;
            const bool has_next = i + 1 < nbt[t];
            const int ph = (cs[t] + i) & 1;
            mbar_wait(bars + S::kPHalf + 2 * t, ph);
            slot_wait(vslot);
            if (i == 0) mbar_wait(bars + S::kOEmpty + t, (cq[t] & 1) ^ 1);  // previous tile's O_t read out
            tc_fence_after();
            pv_half(t, vslot, 0, i > 0);
            mbar_wait(bars + S::kPHalf + 2 * t + 1, ph);
            tc_fence_after();
            pv_half(t, vslot, 1, true);
            if (has_next) {
              slot_wait(knext);
              tc_fence_after();
              qk(t, knext);
              commit(S::kSFull + t);
              if (i + 2 == nbt[t]) commit(S::kQEmpty + t);   // that was the last QK on Q_t
            } else {
              commit(S::kOFull + t);
            }
          }
          commit(S::kKVEmpty + vslot % kNStage);
          if (i + 1 < ti.nblk) commit(S::kKVEmpty + knext % kNStage);
        }
        nbase += 2 * ti.nblk;
        cq[0] += 1; cs[0] += ti.nb0;
        if (ti.act1) { cq[1] += 1; cs[1] += ti.nb1; }
      }
    }
  } else {
    // ================================ softmax warpgroups ===============================
    setmaxnreg_inc<208>();
    const int t = (warp - 4) >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * 128;
    const uint32_t tO = tmem_base + lane_off + 256 + t * D;
    const float sl2 = p.scale_log2;
    const uint64_t sl2x2 = pack_f2(sl2, sl2);
    uint8_t* stage = sStage + t * S::kBoxBytes;
    int cqt = 0, cst = 0;   // tiles / key blocks this warpgroup has processed
    for (int tl = 0;; ++tl) {
      Tile ti;
      if (!next_tile(tl, ti)) break;
      const int nbt = t ? ti.nb1 : ti.nb0;
      if (nbt == 0) continue;   // this tile has no second half
      const int q0 = ti.q0;
      const int QL = ti.QL, KL = ti.KL;
      const int q_row = q0 + t * 128 + row;
      float m_used = -1e30f;
      float l = 0.f;
      for (int i = 0; i < nbt; ++i) {
        mbar_wait(bars + S::kSFull + t, (cst + i) & 1);
        tc_fence_after();
        uint32_t sr[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, sr[c]);
        tmem_ld_wait();
        const int k0 = i * 128;
        const bool need_mask = (k0 + 128 > KL) || (p.causal && (k0 + 127 > q0 + t * 128));
        if (need_mask) {
          const int lim = p.causal ? min(KL - 1, q_row) : (KL - 1);  // last visible key
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (k0 + c * 32 + j > lim) sr[c][j] = 0xff800000u;  // -inf
        }
        auto exp_chunk = [&](int c, float (&pf)[32]) {
          const uint64_t negm2 = pack_f2(-m_used, -m_used);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint64_t x2 = ffma2(pack_f2(__uint_as_float(sr[c][2 * j]), __uint_as_float(sr[c][2 * j + 1])),
                                      sl2x2, negm2);
            float x0, x1;
            unpack_f2(x2, x0, x1);
            pf[2 * j] = fast_exp2(x0);
            pf[2 * j + 1] = fast_exp2(x1);
          }
        };
        float pf[32];
        exp_chunk(0, pf);   // speculative against the running reference max (see the one-tile kernel)
        float mx8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) mx8[u] = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            mx8[j & 7] = fmax3(mx8[j & 7], __uint_as_float(sr[c][2 * j]), __uint_as_float(sr[c][2 * j + 1]));
        const float mx = fmax3(fmax3(mx8[0], mx8[1], mx8[2]), fmax3(mx8[3], mx8[4], mx8[5]),
                               fmaxf(mx8[6], mx8[7]));
        const float mx_s = mx * sl2;
        const bool grow = mx_s > m_used + kRescaleThreshold;
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? mx_s : m_used;
          const float alpha = fast_exp2(m_used - m_new);
          m_used = m_new;
          l *= alpha;
          if (i > 0) {
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
          exp_chunk(0, pf);
        }
        uint64_t sum2 = pack_f2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c > 0) exp_chunk(c, pf);
          uint32_t pr[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            sum2 = fadd2(sum2, pack_f2(pf[2 * j], pf[2 * j + 1]));
            pr[j] = pack2<T>(pf[2 * j], pf[2 * j + 1]);
          }
          tmem_st_x16(tS + c * 16, pr);
          if (c & 1) {  // half of the keys are in TMEM: release the tensor pipe
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + S::kPHalf + 2 * t + (c >> 1));
          }
        }
        float s0, s1;
        unpack_f2(sum2, s0, s1);
        l += s0 + s1;
      }
      cst += nbt;
      // ---- epilogue: O / l -> 16-bit -> private staging, one 64-column box at a time -> TMA store
      mbar_wait(bars + S::kOFull + t, cqt & 1);
      tc_fence_after();
      ++cqt;
      const float inv_l = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
      for (int bx = 0; bx < S::kNBox; ++bx) {
        uint32_t o2[2][32];
        tmem_ld_x32(tO + bx * 64, o2[0]);
        tmem_ld_x32(tO + bx * 64 + 32, o2[1]);
        tmem_ld_wait();
        if (bx == S::kNBox - 1) {   // O_t is in registers: the next tile's first PV may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + S::kOEmpty + t);
        }
        if (wq == 0 && lane == 0) bulk_wait_read<0>();   // the previous store has read the staging box
        named_bar_sync(1 + t, 128);
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 v;
            v.x = pack2<T>(__uint_as_float(o2[c][8 * u + 0]) * inv_l, __uint_as_float(o2[c][8 * u + 1]) * inv_l);
            v.y = pack2<T>(__uint_as_float(o2[c][8 * u + 2]) * inv_l, __uint_as_float(o2[c][8 * u + 3]) * inv_l);
            v.z = pack2<T>(__uint_as_float(o2[c][8 * u + 4]) * inv_l, __uint_as_float(o2[c][8 * u + 5]) * inv_l);
            v.w = pack2<T>(__uint_as_float(o2[c][8 * u + 6]) * inv_l, __uint_as_float(o2[c][8 * u + 7]) * inv_l);
            const int cin = c * 4 + u;  // 16-byte chunk within the 128-byte box row
            *reinterpret_cast<uint4*>(stage + row * 128 + ((cin ^ (row & 7)) << 4)) = v;
          }
        const int rows_left = QL - (q0 + t * 128);  // > 0 here
        if (packed && rows_left < 128) {
          // last, partial tile of a packed sequence: rows past its end belong to the next sequence, so
          // copy the valid rows of this box with 16-byte stores (the barrier that opens the next use of
          // the staging box also closes these reads)
          named_bar_sync(1 + t, 128);
          T* obase = static_cast<T*>(p.o_ptr) +
                     (static_cast<int64_t>(ti.bh_q) * p.total_q + ti.q_off + q0 + t * 128) * D + bx * 64;
          for (int idx = row; idx < rows_left * 8; idx += 128) {
            const int r2 = idx >> 3, cin = idx & 7;
            const uint4 v = *reinterpret_cast<const uint4*>(stage + r2 * 128 + ((cin ^ (r2 & 7)) << 4));
            *reinterpret_cast<uint4*>(obase + static_cast<int64_t>(r2) * D + cin * 8) = v;
          }
        } else {
          fence_proxy_async_smem();
          named_bar_sync(1 + t, 128);
          if (wq == 0 && lane == 0) {
            tma_store_3d(&tm_o, stage, bx * 64, ti.q_off + q0 + t * 128, ti.bh_q);
            bulk_commit();
          }
        }
      }
      if (q_row < QL) {
        const int64_t li = packed ? static_cast<int64_t>(ti.bh_q) * p.total_q + ti.q_off + q_row
                                  : static_cast<int64_t>(ti.bh_q) * QL + q_row;
        p.lse[li] = l > 0.f ? (m_used + fast_log2(l)) * kLn2 : -INFINITY;
      }
    }
    if (wq == 0 && lane == 0) bulk_wait<0>();
  }

  // ---- teardown -------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// x (rows, 64) fp32 -> (rows, 128) fp16 = [hi(64) | lo(64)] of x' = x * 2^-e, hi = fp16(x'), lo = fp16(x' - hi):
// 22 significant bits (bf16 terms would give 16, not enough for 1e-4 absolute on gradients of
// magnitude ~5).  e is the tensor's own binary exponent (|x|max = m * 2^e, m in [0.5, 1); scale block
// below), so x' lies in [-1, 1] whatever the caller's units are: fp16's narrow range (65504 at the top,
// 6e-8 subnormal spacing at the bottom) would otherwise overflow for |x| ~ 1e5 and flush or coarsen a
// tensor of small values -- an upstream gradient dO = 1/N ~ 5e-7 of a mean-reduced loss lost several
// percent.  The power of two is exact; the kernels undo it with the multipliers below.
__global__ void __launch_bounds__(256)
split_f32_kernel(__half* __restrict__ out, const float* __restrict__ in, int64_t n4,
                 const float* __restrict__ scale_slot, int per_row, int half4) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;  // one float4 per thread
  if (i >= n4) return;
  const float sc = scale_slot ? __ldg(scale_slot) : 1.f;   // 2^-e_x: an exact power of two
  float4 x = reinterpret_cast<const float4*>(in)[i];
  x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
  // per_row = E / 4 float4 per E-float row (16 for E = 64); narrower rows leave the rest of each 64-wide
  // half zero, which is what lets E = 16 / 32 ride the same 128-wide kernels
  const int64_t row = i / per_row;
  const int c4 = static_cast<int>(i % per_row);
  const uint32_t h0 = pack2<__half>(x.x, x.y), h1 = pack2<__half>(x.z, x.w);
  const uint32_t l0 = pack2<__half>(x.x - unpack_lo<__half>(h0), x.y - unpack_hi<__half>(h0));
  const uint32_t l1 = pack2<__half>(x.z - unpack_lo<__half>(h1), x.w - unpack_hi<__half>(h1));
  // half4 = width of each half in float4 units: 16 (rows [hi 64 | lo 64]) or 32 (E = 128: [hi 128 | lo 128])
  uint2* o = reinterpret_cast<uint2*>(out + row * (8 * half4));
  o[c4] = make_uint2(h0, h1);
  o[half4 + c4] = make_uint2(l0, l1);
  for (int z = c4 + per_row; z < half4; z += per_row) o[z] = o[half4 + z] = make_uint2(0u, 0u);
}

// split_f32_kernel for up to four tensors and two zero-fills in one launch: blockIdx.y = job
struct StageArgs {
  __half* out[4];
  const float* in[4];
  int64_t n4[4];
  const float* scale[4];
  float4* zero[2];
  int64_t zn4[2];
  int per_row, half4;
};
__global__ void __launch_bounds__(256) f32_stage_kernel(const StageArgs a) {
  const int job = blockIdx.y;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;  // one float4 per thread
  if (job >= 4) {
    if (i < a.zn4[job - 4]) a.zero[job - 4][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  if (i >= a.n4[job]) return;
  const float sc = a.scale[job] ? __ldg(a.scale[job]) : 1.f;
  float4 x = reinterpret_cast<const float4*>(a.in[job])[i];
  x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc;
  const int64_t row = i / a.per_row;
  const int c4 = static_cast<int>(i % a.per_row);
  const uint32_t h0 = pack2<__half>(x.x, x.y), h1 = pack2<__half>(x.z, x.w);
  const uint32_t l0 = pack2<__half>(x.x - unpack_lo<__half>(h0), x.y - unpack_hi<__half>(h0));
  const uint32_t l1 = pack2<__half>(x.z - unpack_lo<__half>(h1), x.w - unpack_hi<__half>(h1));
  uint2* o = reinterpret_cast<uint2*>(a.out[job] + row * (8 * a.half4));
  o[c4] = make_uint2(h0, h1);
  o[a.half4 + c4] = make_uint2(l0, l1);
  for (int z = c4 + a.per_row; z < a.half4; z += a.per_row) o[z] = o[a.half4 + z] = make_uint2(0u, 0u);
}
int launch_split(__half* out, const void* in, int64_t rows, int E, const float* scale_slot, cudaStream_t st) {
  const int64_t n4 = rows * (E / 4);
  if (n4 == 0) return NNOP_OK;
  split_f32_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(out, static_cast<const float*>(in), n4,
                                                                           scale_slot, E / 4, E > 64 ? 32 : 16);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

// ---- Float32 scale block (kF32ScaleBytes at the end of the Float32 workspaces) ----
//   u32[0..3]   bit patterns of |q|max, |k|max, |v|max, |dO|max (atomicMax; non-negative floats order as uints)
//   i32[4..7]   their binary exponents e_q, e_k, e_v, e_dO (0 for an all-zero or non-finite tensor), clamped to
//               [-120, 120] so that 2^-e stays a normal float
//   f32[8..14]  multipliers, see F32Mult in internal.h
//   f32[16..19] 2^-e_x, what the split kernels multiply the inputs by
//   u32[20]     blocks of the |x|max pass that have finished: the last one writes [4..19] (no extra launch)
struct AbsmaxArgs {
  const float* ptr[4];
  int64_t n4[4];
};
__global__ void __launch_bounds__(256) absmax_f32_kernel(uint32_t* __restrict__ blk, const AbsmaxArgs a) {
  const int which = blockIdx.y;
  const float4* in = reinterpret_cast<const float4*>(a.ptr[which]);
  const int64_t n4 = a.n4[which];
  float m = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
  int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {   // four independent loads in flight per thread
    const float4 x0 = in[i], x1 = in[i + stride], x2 = in[i + 2 * stride], x3 = in[i + 3 * stride];
    const float m0 = fmaxf(fmaxf(fabsf(x0.x), fabsf(x0.y)), fmaxf(fabsf(x0.z), fabsf(x0.w)));
    const float m1 = fmaxf(fmaxf(fabsf(x1.x), fabsf(x1.y)), fmaxf(fabsf(x1.z), fabsf(x1.w)));
    const float m2 = fmaxf(fmaxf(fabsf(x2.x), fabsf(x2.y)), fmaxf(fabsf(x2.z), fabsf(x2.w)));
    const float m3 = fmaxf(fmaxf(fabsf(x3.x), fabsf(x3.y)), fmaxf(fabsf(x3.z), fabsf(x3.w)));
    m = fmaxf(m, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
  }
  for (; i < n4; i += stride) {
    const float4 x = in[i];
    m = fmaxf(fmaxf(m, fmaxf(fabsf(x.x), fabsf(x.y))), fmaxf(fabsf(x.z), fabsf(x.w)));
  }
  m = warp_max(m);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x != 0) return;
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  if (m > 0.f) atomicMax(blk + which, __float_as_uint(m));
  __threadfence();
  if (atomicAdd(blk + 20, 1u) != gridDim.x * gridDim.y - 1) return;
  // last block: exponents, input scales and the multipliers that undo them
  __threadfence();
  int e[4];
  float* f = reinterpret_cast<float*>(blk);
  for (int i = 0; i < 4; ++i) {
    const float mx = __uint_as_float(atomicOr(blk + i, 0u));   // (atomic read: the other blocks' maxima)
    int ex = 0;
    if (mx > 0.f && mx <= 3.4028234e38f) frexpf(mx, &ex);
    ex = ex < -120 ? -120 : (ex > 120 ? 120 : ex);
    e[i] = ex;
    reinterpret_cast<int*>(blk)[4 + i] = ex;
    f[16 + i] = ldexpf(1.f, -ex);
  }
  const int eq = e[0], ek = e[1], ev = e[2], edo = e[3];
  f[8 + F32Mult::kLogits] = ldexpf(1.f, eq + ek);
  f[8 + F32Mult::kO] = ldexpf(1.f, ev);
  f[8 + F32Mult::kDeltaInv] = ldexpf(1.f, -(edo + ev));
  f[8 + F32Mult::kDV] = ldexpf(1.f, edo);
  f[8 + F32Mult::kDQ] = ldexpf(1.f, edo + ev + ek);
  f[8 + F32Mult::kDK] = ldexpf(1.f, edo + ev + eq);
  f[8 + F32Mult::kDPair] = ldexpf(1.f, edo + ev);
}

// Float32, E = 64: split q, k, v into [hi | lo] bf16 rows in the workspace, then the SPLIT kernel
template <bool BIAS, int D = 128>
int launch_fwd_f32(const AttnParams& a) {
  using T = __half;
  using S = FwdSmem<D, D == 256 ? 2 : (BIAS ? 3 : 4), BIAS ? 65536 : 0>;
  constexpr int kCtaRows = D == 256 ? 128 : 256;
  const int64_t rq = static_cast<int64_t>(a.B) * a.QH * a.QL, rk = static_cast<int64_t>(a.B) * a.KH * a.KL;
  T* qs = static_cast<T*>(a.fwd_ws);
  T* ks = qs + rq * D;
  T* vs = ks + rk * D;
  void* blk = vs + rk * D;   // scale block (256-byte aligned: every copy is a multiple of 256 bytes)
  if (int rc = attn_f32_scales(blk, a.q, rq * a.E, a.k, rk * a.E, a.v, rk * a.E, nullptr, 0, a.stream)) return rc;
  {
    F32StageJobs jobs;   // q, k, v -> [hi | lo] in one launch
    jobs.out[0] = qs; jobs.in[0] = a.q; jobs.rows[0] = rq; jobs.scale[0] = f32_in_scale(blk, 0);
    jobs.out[1] = ks; jobs.in[1] = a.k; jobs.rows[1] = rk; jobs.scale[1] = f32_in_scale(blk, 1);
    jobs.out[2] = vs; jobs.in[2] = a.v; jobs.rows[2] = rk; jobs.scale[2] = f32_in_scale(blk, 2);
    if (int rc = attn_stage_f32(jobs, a.E, a.stream)) return rc;
  }
  alignas(64) CUtensorMap tq, tk, tv, to;
  const uint64_t bhq = static_cast<uint64_t>(a.B) * a.QH, bhk = static_cast<uint64_t>(a.B) * a.KH;
  if (int rc = make_tmap_3d(&tq, qs, NNOP_F16, D, a.QL, bhq, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tk, ks, NNOP_F16, D, a.KL, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tv, vs, NNOP_F16, D, a.KL, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&to, a.o, NNOP_F32, a.E, a.QL, bhq, 32, 128)) return rc;   // E < 64: columns >= E clipped
  alignas(64) CUtensorMap tb = to;  // unused without a bias
  if constexpr (BIAS)
    if (int rc = make_tmap_3d(&tb, a.pair_t, NNOP_F32, a.KLp, a.QL, bhq, 32, 128)) return rc;
  auto kern = attn_fwd_sm100_kernel<T, D, true, BIAS>;
  NNOP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kDynBytes));
  FwdParams fp;
  fp.lse = a.lse;
  fp.QL = a.QL; fp.KL = a.KL; fp.QH = a.QH; fp.KH = a.KH; fp.causal = a.causal;
  fp.scale_log2 = a.scale * kLog2e;
  fp.f32_mult = f32_mults(blk);
  fp.cu_q = nullptr; fp.cu_k = nullptr; fp.o_ptr = a.o; fp.total_q = 0; fp.nseq = 0;
  fp.o_row_bytes = a.E * static_cast<int>(sizeof(float));
  fp.lpt_group = 0;
  fp.kpad = a.kpad;
  dim3 grid((a.QL + kCtaRows - 1) / kCtaRows, a.QH, a.B);
  timing_begin(0, a.stream);
  kern<<<grid, kFwdThreads, S::kDynBytes, a.stream>>>(tq, tk, tv, to, tb, fp);
  timing_end(0, a.stream);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

template <typename T, int D, bool BIAS = false, bool QUAD = false>
int launch_fwd(const AttnParams& a) {
  using S = FwdSmem<D, D == 256 ? 2 : (BIAS ? 3 : 4), BIAS ? 65536 : 0>;
  constexpr int kCtaRows = D == 256 ? 128 : 256;
  alignas(64) CUtensorMap tq, tk, tv, to;
  const bool packed = a.cu_q != nullptr;
  // packed mode: the tensors are (QH, total_q, E) / (KH, total_k, E); a.QL / a.KL hold the maxima
  const uint64_t bhq = packed ? a.QH : static_cast<uint64_t>(a.B) * a.QH;
  const uint64_t bhk = packed ? a.KH : static_cast<uint64_t>(a.B) * a.KH;
  const uint64_t rows_q = packed ? static_cast<uint64_t>(a.total_q) : a.QL;
  const uint64_t rows_k = packed ? static_cast<uint64_t>(a.total_k) : a.KL;
  // a.E < D (embedding dims 16 / 32 on the D = 64 kernels): the maps describe the real E-wide rows, the
  // boxes stay 64 wide -- TMA zero-fills columns >= E on the way in and clips them on the way out
  if (int rc = make_tmap_3d(&tq, a.q, a.dtype, a.E, rows_q, bhq, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tk, a.k, a.dtype, a.E, rows_k, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tv, a.v, a.dtype, a.E, rows_k, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&to, a.o, a.dtype, a.E, rows_q, bhq, 64, 128)) return rc;
  alignas(64) CUtensorMap tb = to;  // unused without a bias
  if constexpr (BIAS)
    if (int rc = make_tmap_3d(&tb, a.pair_t, a.dtype, a.KLp, a.QL, bhq, 64, 128)) return rc;
  auto kern = attn_fwd_sm100_kernel<T, D, false, BIAS, QUAD>;
  constexpr int kSmemBytes = S::kDynBytes + (QUAD ? 16 + kFwdQuadXchBytes : 0);
  NNOP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  FwdParams fp;
  fp.lse = a.lse;
  fp.QL = a.QL; fp.KL = a.KL; fp.QH = a.QH; fp.KH = a.KH; fp.causal = a.causal;
  fp.scale_log2 = a.scale * kLog2e;
  fp.f32_mult = nullptr;
  fp.cu_q = a.cu_q; fp.cu_k = a.cu_k; fp.o_ptr = a.o; fp.total_q = a.total_q;
  fp.o_row_bytes = a.E * static_cast<int>(sizeof(T));
  fp.kpad = packed ? nullptr : a.kpad;
  fp.nseq = a.nseq;
  fp.lpt_group = 0;
  if (!packed && a.causal) {
    // (batch, head) pairs whose K + V stay under ~48 MB of L2 together; GQA heads share theirs
    const double kv_bytes = 2.0 * a.KL * D * sizeof(T) * (static_cast<double>(a.KH) / a.QH);
    const int env = getenv("NNOP_FWD_LPT_GROUP") ? atoi(getenv("NNOP_FWD_LPT_GROUP")) : -1;
    fp.lpt_group = env >= 0 ? env : static_cast<int>(48.0 * 1024 * 1024 / (kv_bytes > 1 ? kv_bytes : 1));
    if (fp.lpt_group > a.QH * a.B) fp.lpt_group = a.QH * a.B;
  }
  dim3 grid(packed ? static_cast<unsigned>(a.total_q / 256 + a.nseq) : (a.QL + kCtaRows - 1) / kCtaRows, a.QH,
            packed ? 1 : a.B);
  timing_begin(0, a.stream);
  kern<<<grid, QUAD ? kFwdQuadThreads : kFwdThreads, kSmemBytes, a.stream>>>(tq, tk, tv, to, tb, fp);
  timing_end(0, a.stream);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

std::atomic<int> g_fwd_mode{-1};

inline int fwd_mode() {
  int m = g_fwd_mode.load();
  if (m < 0) {
    const char* e = getenv("NNOP_FWD_MODE");
    m = e ? atoi(e) : 0;
    g_fwd_mode.store(m);
  }
  return m;
}

template <typename T, int D>
int launch_fwd_persist(const AttnParams& a, int ctas) {
  using S = FwdPersistSmem<D>;
  alignas(64) CUtensorMap tq, tk, tv, to;
  const bool packed = a.cu_q != nullptr;
  const uint64_t bhq = packed ? a.QH : static_cast<uint64_t>(a.B) * a.QH;
  const uint64_t bhk = packed ? a.KH : static_cast<uint64_t>(a.B) * a.KH;
  const uint64_t rows_q = packed ? static_cast<uint64_t>(a.total_q) : a.QL;
  const uint64_t rows_k = packed ? static_cast<uint64_t>(a.total_k) : a.KL;
  // a.E < D (embedding dims 16 / 32 on the D = 64 kernels): the maps describe the real E-wide rows, the
  // boxes stay 64 wide -- TMA zero-fills columns >= E on the way in and clips them on the way out
  if (int rc = make_tmap_3d(&tq, a.q, a.dtype, a.E, rows_q, bhq, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tk, a.k, a.dtype, a.E, rows_k, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tv, a.v, a.dtype, a.E, rows_k, bhk, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&to, a.o, a.dtype, a.E, rows_q, bhq, 64, 128)) return rc;
  auto kern = attn_fwd_sm100_persist_kernel<T, D>;
  NNOP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kDynBytes));
  FwdParams fp;
  fp.lse = a.lse;
  fp.QL = a.QL; fp.KL = a.KL; fp.QH = a.QH; fp.KH = a.KH; fp.causal = a.causal;
  fp.scale_log2 = a.scale * kLog2e;
  fp.f32_mult = nullptr;
  fp.cu_q = a.cu_q; fp.cu_k = a.cu_k; fp.o_ptr = a.o; fp.total_q = a.total_q; fp.kpad = nullptr;
  fp.o_row_bytes = a.E * static_cast<int>(sizeof(T));
  fp.nseq = a.nseq;
  fp.lpt_group = 0;
  // the tile counter lives in the caller's workspace (the library keeps no device state of its own):
  // zeroed on the call's stream, so concurrent launches and replayed graphs each use their own
  int* counter = static_cast<int*>(a.fwd_ws);
  NNOP_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), a.stream));
  const int nqt = (a.QL + 255) / 256;
  // packed: the exact tile count is only known on the device; this is its upper bound (for the grid)
  const int64_t n_tiles = packed ? (a.total_q / 256 + a.nseq) * a.QH : static_cast<int64_t>(nqt) * a.QH * a.B;
  const int grid = static_cast<int>(n_tiles < ctas ? n_tiles : ctas);
  timing_begin(0, a.stream);
  kern<<<grid, kFwdThreads, S::kDynBytes, a.stream>>>(tq, tk, tv, to, fp, counter, static_cast<int>(n_tiles), nqt);
  timing_end(0, a.stream);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

}  // namespace

void attn_sm100_set_fwd_mode(int mode) { g_fwd_mode.store(mode); }

#ifdef NNOP_FWD_TRACE
extern "C" int nnop_debug_fwd_trace(long long* host_out, int n) {
  return static_cast<int>(cudaMemcpyFromSymbol(host_out, g_fwd_trace, sizeof(long long) * n));
}
#endif

bool attn_sm100_supported(const AttnParams& a, bool backward) {
  if (backward && !attn_sm100_bwd_available()) return false;
  // Float32 forward: E <= 64 on the two-tile split kernel (also with a bias), E = 128 on the one-tile form (no bias)
  const bool f32_split = a.dtype == NNOP_F32 && !backward && a.fwd_ws != nullptr && !a.cu_q &&
                         (a.E == 16 || a.E == 32 || a.E == 64 || (a.E == 128 && a.pair == nullptr));
  if (a.dtype != NNOP_F16 && a.dtype != NNOP_BF16 && !f32_split) return false;
  if (a.E != 64 && a.E != 128) {
    // 16 / 32: dense 16-bit problems only (the packed kernels address partial tiles with plain pointers)
    // 256: dense 16-bit FORWARD without a bias (one q tile per CTA); its backward stays on the SIMT kernels --
    // dK and dV alone would fill all 512 TMEM columns
    const bool e256 = a.E == 256 && !backward && a.dtype != NNOP_F32 && a.pair == nullptr;
    if ((a.E != 16 && a.E != 32 && !e256) || a.cu_q != nullptr) return false;
  }
  // the additive bias needs its head-major copy (and, backward, a dpair staging area) in the
  // workspace (nnop_flash_attn_pair_workspace_bytes); without it the generic path serves it
  if (a.pair && (a.pair_t == nullptr || a.cu_q != nullptr)) return false;
  if (a.pair && backward && a.dpair_t == nullptr) return false;
  if (a.QL < 1 || a.KL < 1) return false;
  if (a.QH > 65535 || a.B > 65535 || a.nseq > 65535) return false;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al(a.q) || !al(a.k) || !al(a.v) || !al(a.o)) return false;
  if (backward && (!al(a.dq) || !al(a.dk) || !al(a.dv) || !al(a.dO))) return false;
  return true;
}

int attn_split_f32_rows(void* out_bf16x2, const void* in_f32, int64_t rows, int E, const float* scale_slot,
                        cudaStream_t st) {
  return launch_split(static_cast<__half*>(out_bf16x2), in_f32, rows, E, scale_slot, st);
}

int attn_stage_f32(const F32StageJobs& jobs, int E, cudaStream_t st) {
  StageArgs a;
  int64_t most = 0;
  int last = -1;
  for (int j = 0; j < 4; ++j) {
    a.out[j] = static_cast<__half*>(jobs.out[j]);
    a.in[j] = static_cast<const float*>(jobs.in[j]);
    a.n4[j] = jobs.out[j] ? jobs.rows[j] * (E / 4) : 0;
    a.scale[j] = jobs.scale[j];
    if (a.n4[j] > 0) last = j;
    if (a.n4[j] > most) most = a.n4[j];
  }
  for (int j = 0; j < 2; ++j) {
    a.zero[j] = static_cast<float4*>(jobs.zero[j]);
    a.zn4[j] = jobs.zero[j] ? jobs.zero_floats[j] / 4 : 0;
    if (a.zn4[j] > 0) last = 4 + j;
    if (a.zn4[j] > most) most = a.zn4[j];
  }
  if (last < 0) return NNOP_OK;
  a.per_row = E / 4;
  a.half4 = E > 64 ? 32 : 16;
  f32_stage_kernel<<<dim3(static_cast<unsigned>((most + 255) / 256), last + 1), 256, 0, st>>>(a);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

int attn_f32_scales(void* block, const void* q, int64_t nq, const void* k, int64_t nk, const void* v,
                    int64_t nv, const void* dO, int64_t ndo, cudaStream_t st) {
  NNOP_CUDA_CHECK(cudaMemsetAsync(block, 0, kF32ScaleBytes, st));
  AbsmaxArgs a;
  const void* ptrs[4] = {q, k, v, dO};
  const int64_t ns[4] = {nq, nk, nv, dO ? ndo : 0};
  int64_t most = 0;
  for (int i = 0; i < 4; ++i) {
    a.ptr[i] = static_cast<const float*>(ptrs[i]);
    a.n4[i] = ns[i] / 4;   // E = 64: element counts are multiples of 64
    if (a.n4[i] > most) most = a.n4[i];
  }
  int64_t gx = (most + 256 * 8 - 1) / (256 * 8);   // ~8 float4 per thread
  if (gx < 1) gx = 1;
  if (gx > 4 * sm_count()) gx = 4 * sm_count();
  absmax_f32_kernel<<<dim3(static_cast<unsigned>(gx), dO ? 4 : 3), 256, 0, st>>>(static_cast<uint32_t*>(block), a);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

size_t attn_sm100_fwd_workspace_bytes(int dtype, int E, int QL, int KL, int QH, int KH, int B) {
  // 16-bit: the persistent forward's tile counter; Float32 E = 64: the [hi | lo] copies + scale block
  if (dtype != NNOP_F32) return (E == 16 || E == 32 || E == 64 || E == 128) ? kFwdCounterBytes : 0;
  if (E != 16 && E != 32 && E != 64 && E != 128) return 0;
  const size_t row = E == 128 ? 256 : 128;   // [hi | lo] fp16 elements per row
  return (static_cast<size_t>(B) * QH * QL + 2 * static_cast<size_t>(B) * KH * KL) * row * 2 + kF32ScaleBytes;
}

int attn_sm100_fwd(const AttnParams& a) {
  if (a.pair) {
    if (int rc = attn_pair_to_head_major(a)) return rc;
    if (a.dtype == NNOP_F32) return launch_fwd_f32<true>(a);   // (E <= 64: attn_sm100_supported)
    if (a.dtype == NNOP_BF16)
      return a.E == 128 ? launch_fwd<__nv_bfloat16, 128, true>(a) : launch_fwd<__nv_bfloat16, 64, true>(a);
    return a.E == 128 ? launch_fwd<__half, 128, true>(a) : launch_fwd<__half, 64, true>(a);
  }
  if (a.dtype == NNOP_F32) return a.E == 128 ? launch_fwd_f32<false, 256>(a) : launch_fwd_f32<false>(a);
  if (a.E == 256) return a.dtype == NNOP_BF16 ? launch_fwd<__nv_bfloat16, 256>(a) : launch_fwd<__half, 256>(a);
  // forward variant (nnop_set_fwd_mode / NNOP_FWD_MODE): 0 automatic, 1 one CTA per q tile, 2 persistent,
  // 100+n persistent on n CTAs
  const int mode = fwd_mode();
  const bool packed = a.cu_q != nullptr;
  const int64_t n_tiles = packed ? (a.total_q / 256 + a.nseq) * a.QH
                                 : static_cast<int64_t>((a.QL + 255) / 256) * a.QH * a.B;
  const bool persist_ok = a.fwd_ws != nullptr && a.kpad == nullptr && (packed || a.KL >= 1) && n_tiles < (1LL << 30);
  // automatic choice, from measurement (profiles/r01d_perf_fwd_modes.txt): the persistent kernel wins
  // where the per-tile fixed cost matters -- E = 64 (+15 %) and short sequences (L = 2 048: +6 %) --
  // and is neutral (bench.py regime) to slower (back-to-back launches) on long E = 128 tiles
  // (packed batches: sequences of very different lengths -- the dynamic queue balances them; C4 1 038 -> 1 067 TFLOP/s)
  const bool persist_pays = a.E <= 64 || a.QL <= 2048 || packed;
  if (persist_ok && (mode == 2 || mode >= 100 || (mode == 0 && persist_pays && n_tiles >= 2LL * sm_count()))) {
    const int ctas = mode >= 100 ? (mode - 100 < 1 ? 1 : mode - 100) : sm_count();
    if (a.dtype == NNOP_BF16)
      return a.E == 128 ? launch_fwd_persist<__nv_bfloat16, 128>(a, ctas) : launch_fwd_persist<__nv_bfloat16, 64>(a, ctas);
    return a.E == 128 ? launch_fwd_persist<__half, 128>(a, ctas) : launch_fwd_persist<__half, 64>(a, ctas);
  }
  // 3 = two softmax warps per 32 rows (QUAD; E = 128, dense layout)
  if (mode == 3 && a.E == 128 && !packed)
    return a.dtype == NNOP_BF16 ? launch_fwd<__nv_bfloat16, 128, false, true>(a) : launch_fwd<__half, 128, false, true>(a);
  if (a.dtype == NNOP_BF16)
    return a.E == 128 ? launch_fwd<__nv_bfloat16, 128>(a) : launch_fwd<__nv_bfloat16, 64>(a);
  return a.E == 128 ? launch_fwd<__half, 128>(a) : launch_fwd<__half, 64>(a);
}

}  // namespace nnop
