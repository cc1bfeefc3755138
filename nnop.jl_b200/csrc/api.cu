// api.cu -- C-ABI entry points of libnnop_b200.so for flash attention, plus error / device
// plumbing.  Validation mirrors the reference launcher (`_flash_attention`,
// src/attention.jl:141-144; `∇flash_attention`, src/attention_bwd.jl:210-213), including the
// wording of its error strings.
#include <cuda.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include "internal.h"

namespace nnop {

static thread_local char g_err[512] = "";
static thread_local int g_last_path = 0;
static std::atomic<int> g_path_mode{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void clear_error() { g_err[0] = '\0'; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn lookup_encode_fn() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    return reinterpret_cast<EncodeTiledFn>(p);
  return nullptr;
}
static EncodeTiledFn get_encode_fn() {
  static const EncodeTiledFn fn = lookup_encode_fn();   // initialised once, thread-safe (C++11 static)
  return fn;
}

int make_tmap_3d(void* tmap_out, const void* base, int dtype, uint64_t inner, uint64_t rows,
                 uint64_t outer, uint32_t box_inner, uint32_t box_rows) {
  const uint64_t elem_bytes = dtype == NNOP_F32 ? 4 : 2;
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(NNOP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[3] = {inner, rows, outer};
  const cuuint64_t strides[2] = {inner * elem_bytes, inner * rows * elem_bytes};
  const cuuint32_t box[3] = {box_inner, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = dtype == NNOP_F32   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : dtype == NNOP_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                     : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(static_cast<CUtensorMap*>(tmap_out), dt, 3, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(NNOP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return NNOP_OK;
}

static thread_local cudaEvent_t g_ev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
void timing_begin(int which, cudaStream_t st) {
  if (g_ev[which][0]) cudaEventRecord(g_ev[which][0], st);
}
void timing_end(int which, cudaStream_t st) {
  if (g_ev[which][1]) cudaEventRecord(g_ev[which][1], st);
  g_ev[which][0] = g_ev[which][1] = nullptr;
}

static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

static int validate(int dtype, int E, int QL, int KL, int QH, int KH, int B) {
  if (dtype != NNOP_F32 && dtype != NNOP_F16 && dtype != NNOP_BF16)
    return fail(NNOP_ERR_DTYPE, "unknown dtype code %d", dtype);
  if (E <= 0 || QL < 0 || KL < 0 || QH <= 0 || KH <= 0 || B < 0)
    return fail(NNOP_ERR_SHAPE, "Invalid attention shape E=%d QL=%d KL=%d QH=%d KH=%d B=%d.", E, QL,
                KL, QH, KH, B);
  if (!is_pow2(E)) return fail(NNOP_ERR_UNSUPPORTED_E, "Only power-of-2 embedding dims are supported.");
  if (E < 16 || E > 256)
    return fail(NNOP_ERR_UNSUPPORTED_E,
                "Embedding dim `%d` is not supported (power of 2 in [16, 256]).", E);
  if (QH % KH != 0)
    return fail(NNOP_ERR_SHAPE,
                "Number of query heads `%d` must be divisible by number of KV heads `%d`.", QH, KH);
  return NNOP_OK;
}

}  // namespace nnop

using namespace nnop;

extern "C" int nnop_version(void) { return NNOP_B200_VERSION; }
extern "C" const char* nnop_last_error_string(void) { return g_err; }

extern "C" int nnop_device_info(int device, nnop_device_info_t* out) {
  clear_error();
  if (!out) return fail(NNOP_ERR_ARG, "NULL pointer");
  cudaDeviceProp prop;
  NNOP_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  out->sm_count = prop.multiProcessorCount;
  out->cc_major = prop.major;
  out->cc_minor = prop.minor;
  out->shared_mem_per_block_optin = prop.sharedMemPerBlockOptin;
  out->l2_bytes = static_cast<size_t>(prop.l2CacheSize);
  out->hbm_bytes = prop.totalGlobalMem;
  return NNOP_OK;
}

extern "C" int nnop_set_timing_events(int which, void* start_event, void* stop_event) {
  if (which < 0 || which > 1) return fail(NNOP_ERR_ARG, "which must be 0 (forward) or 1 (backward)");
  g_ev[which][0] = static_cast<cudaEvent_t>(start_event);
  g_ev[which][1] = static_cast<cudaEvent_t>(stop_event);
  return NNOP_OK;
}

extern "C" int nnop_set_attention_path(int mode) {
  if (mode < 0 || mode > 2) return fail(NNOP_ERR_ARG, "attention path mode must be 0, 1 or 2");
  g_path_mode.store(mode);
  return NNOP_OK;
}
extern "C" int nnop_last_attention_path(void) { return g_last_path; }
extern "C" int nnop_set_fwd_mode(int mode) {
  if (mode < 0 || (mode > 3 && mode < 101) || mode > 100 + 65535)
    return fail(NNOP_ERR_ARG, "forward kernel mode must be 0..3 or 100+n");
  attn_sm100_set_fwd_mode(mode);
  return NNOP_OK;
}
extern "C" int nnop_set_bwd_pair_mode(int mode) {
  if (mode < 0 || (mode > 4 && mode < 101) || mode > 100 + 65535)
    return fail(NNOP_ERR_ARG, "backward kernel mode must be 0..4 or 100+n");
  attn_sm100_set_bwd_pair_mode(mode);
  return NNOP_OK;
}

extern "C" size_t nnop_flash_attn_fwd_workspace_bytes(int dtype, int E, int QL, int KL, int QH,
                                                      int KH, int B) {
  if (E <= 0 || QL <= 0 || KL <= 0 || QH <= 0 || KH <= 0 || B <= 0) return 0;
  return attn_sm100_fwd_workspace_bytes(dtype, E, QL, KL, QH, KH, B);
}

extern "C" size_t nnop_flash_attn_pair_workspace_bytes(int dtype, int QL, int KL, int QH, int B,
                                                       int backward) {
  if (dtype != NNOP_F32 && dtype != NNOP_F16 && dtype != NNOP_BF16) return 0;
  return attn_pair_workspace_bytes(dtype, QL, KL, QH, B, backward != 0);
}

static inline size_t up256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

extern "C" int nnop_flash_attn_fwd(void* o, float* lse, const void* q, const void* k, const void* v,
                                   const void* pair, const uint8_t* kpad_mask, int dtype, int E,
                                   int QL, int KL, int QH, int KH, int B, int causal, float scale,
                                   void* stream) {
  return nnop_flash_attn_fwd_ws(o, lse, q, k, v, pair, kpad_mask, dtype, E, QL, KL, QH, KH, B, causal,
                                scale, nullptr, 0, stream);
}

extern "C" int nnop_flash_attn_fwd_ws(void* o, float* lse, const void* q, const void* k,
                                      const void* v, const void* pair, const uint8_t* kpad_mask,
                                      int dtype, int E, int QL, int KL, int QH, int KH, int B,
                                      int causal, float scale, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  clear_error();
  if (int rc = validate(dtype, E, QL, KL, QH, KH, B)) return rc;
  if (static_cast<int64_t>(B) * QL == 0) return NNOP_OK;
  if (!o || !lse || !q || (KL > 0 && (!k || !v))) return fail(NNOP_ERR_ARG, "NULL pointer");
  AttnParams p{};
  p.o = o; p.lse = lse; p.q = q; p.k = k; p.v = v; p.pair = pair; p.kpad = kpad_mask;
  p.dtype = dtype; p.E = E; p.QL = QL; p.KL = KL; p.QH = QH; p.KH = KH; p.B = B;
  p.causal = causal ? 1 : 0; p.scale = scale;
  p.stream = static_cast<cudaStream_t>(stream);
  if (workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && KL > 0 &&
      workspace_bytes >= attn_sm100_fwd_workspace_bytes(dtype, E, QL, KL, QH, KH, B) &&
      attn_sm100_fwd_workspace_bytes(dtype, E, QL, KL, QH, KH, B) > 0)
    p.fwd_ws = workspace;  // enables the tensor-core Float32 forward (E = 64)
  if (pair && workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && KL > 0) {
    // pair bias on the tensor cores: its head-major copy lives behind the base workspace
    const size_t base = up256(attn_sm100_fwd_workspace_bytes(dtype, E, QL, KL, QH, KH, B));
    if (workspace_bytes >= base + attn_pair_workspace_bytes(dtype, QL, KL, QH, B, false)) {
      p.pair_t = static_cast<char*>(workspace) + base;
      p.KLp = pair_klp(KL);
    }
  }
  const int mode = g_path_mode.load();
  const bool fast_ok = attn_sm100_supported(p, false);
  if (mode == 2 && !fast_ok)
    return fail(NNOP_ERR_ARG, "tcgen05 attention path required but the problem does not qualify");
  if (fast_ok && mode != 1) {
    g_last_path = 1;
    return attn_sm100_fwd(p);
  }
  g_last_path = 0;
  return attn_generic_fwd(p);
}

extern "C" size_t nnop_flash_attn_bwd_workspace_bytes(int dtype, int E, int QL, int KL, int QH,
                                                      int KH, int B) {
  if (E <= 0 || QL <= 0 || QH <= 0 || B <= 0) return 0;
  // delta (B,QH,QL) fp32, 256-byte aligned, then the fp32 dQ accumulator of the tcgen05 path
  size_t delta = (static_cast<size_t>(B) * QH * QL * sizeof(float) + 255) & ~static_cast<size_t>(255);
  size_t rest = attn_sm100_bwd_workspace_bytes(E, QL, QH, B);
  if (dtype == NNOP_F32 && (E == 16 || E == 32 || E == 64) && KL > 0 && KH > 0) {  // split-bf16 copies of q, k, v, dO (tensor-core Float32 path)
    const size_t f32 = attn_f32_bwd_workspace_bytes(QL, KL, QH, KH, B);
    if (f32 > rest) rest = f32;
  }
  return delta + rest;
}

static int flash_attn_bwd_impl(void* dq, void* dk, void* dv, void* dpair, const void* dO,
                               const void* o, const float* lse, const void* q, const void* k,
                               const void* v, const void* pair, const uint8_t* kpad_mask,
                               int dtype, int E, int QL, int KL, int QH, int KH, int B,
                               int causal, float scale, void* workspace, size_t workspace_bytes,
                               void* stream, const void* pair_head_major) {
  clear_error();
  if (int rc = validate(dtype, E, QL, KL, QH, KH, B)) return rc;
  if (static_cast<int64_t>(B) * QL == 0 && static_cast<int64_t>(B) * KL == 0) return NNOP_OK;
  if ((QL > 0 && (!dq || !dO || !o || !lse || !q)) || (KL > 0 && (!dk || !dv || !k || !v)))
    return fail(NNOP_ERR_ARG, "NULL pointer");
  if (pair && !dpair) return fail(NNOP_ERR_ARG, "dpair must be given when pair is given");
  const size_t need = nnop_flash_attn_bwd_workspace_bytes(dtype, E, QL, KL, QH, KH, B);
  if (need > 0 && (!workspace || workspace_bytes < need))
    return fail(NNOP_ERR_WORKSPACE, "flash attention backward needs a %zu-byte workspace, got %zu",
                need, workspace_bytes);
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return fail(NNOP_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  AttnParams p{};
  p.o = const_cast<void*>(o); p.lse = const_cast<float*>(lse);
  p.q = q; p.k = k; p.v = v; p.pair = pair; p.kpad = kpad_mask;
  p.dq = dq; p.dk = dk; p.dv = dv; p.dpair = pair ? dpair : nullptr; p.dO = dO;
  p.delta = static_cast<float*>(workspace);
  const size_t delta_bytes =
      (static_cast<size_t>(B) * QH * QL * sizeof(float) + 255) & ~static_cast<size_t>(255);
  p.dq_accum = reinterpret_cast<float*>(static_cast<char*>(workspace) + delta_bytes);
  p.dtype = dtype; p.E = E; p.QL = QL; p.KL = KL; p.QH = QH; p.KH = KH; p.B = B;
  p.causal = causal ? 1 : 0; p.scale = scale;
  p.stream = static_cast<cudaStream_t>(stream);
  if (pair && KL > 0 && QL > 0) {
    // pair bias on the tensor cores: head-major pair and dpair staging behind the base workspace
    const size_t base = up256(need);
    const size_t one = attn_pair_workspace_bytes(dtype, QL, KL, QH, B, false);
    if (pair_head_major && (reinterpret_cast<uintptr_t>(pair_head_major) & 255) == 0 &&
        workspace_bytes >= base + one) {
      // the forward's head-major copy is still alive: no second layout change of pair
      p.pair_t = const_cast<void*>(pair_head_major);
      p.pair_t_ready = 1;
      p.dpair_t = static_cast<char*>(workspace) + base;
      p.KLp = pair_klp(KL);
    } else if (workspace_bytes >= base + 2 * one) {
      p.pair_t = static_cast<char*>(workspace) + base;
      p.dpair_t = static_cast<char*>(workspace) + base + one;
      p.KLp = pair_klp(KL);
    }
  }
  const int mode = g_path_mode.load();
  // Float32, E = 64: split-fp16 tensor-core backward (attn_bwd_f32_sm100.cu)
  const bool al16 = ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) |
                      reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(dO) |
                      reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) |
                      reinterpret_cast<uintptr_t>(dv)) & 15) == 0;
  const bool f32_tc = dtype == NNOP_F32 && (E == 16 || E == 32 || E == 64) && (!pair || p.pair_t) && QL > 0 && KL > 0 && al16 &&
                      QH <= 65535 && B <= 65535;
  const bool fast_ok = f32_tc || attn_sm100_supported(p, true);
  if (mode == 2 && !fast_ok)
    return fail(NNOP_ERR_ARG, "tcgen05 attention path required but the problem does not qualify");
  if (fast_ok && mode != 1) {
    g_last_path = 1;
    if (f32_tc) return attn_f32_bwd(p);
    return attn_sm100_bwd(p);  // runs its own preprocess (delta, lse2, dQ accumulator zeroing)
  }
  g_last_path = 0;
  if (QL > 0) {
    if (int rc = attn_bwd_preprocess(p)) return rc;
  }
  if (QL == 0 || KL == 0) {
    // degenerate: gradients are all zero
    if (KL > 0) {
      const size_t kvb = static_cast<size_t>(B) * KH * KL * E * dtype_size(dtype);
      NNOP_CUDA_CHECK(cudaMemsetAsync(dk, 0, kvb, p.stream));
      NNOP_CUDA_CHECK(cudaMemsetAsync(dv, 0, kvb, p.stream));
    }
    if (QL > 0)
      NNOP_CUDA_CHECK(cudaMemsetAsync(dq, 0, static_cast<size_t>(B) * QH * QL * E * dtype_size(dtype),
                                      p.stream));
    return NNOP_OK;
  }
  return attn_generic_bwd(p);
}

// ---------------------------------------------------------------------------------------
// packed variable-length attention (additive API): tcgen05 path only
// ---------------------------------------------------------------------------------------
static int validate_varlen(int dtype, int E, int nseq, int max_q, int max_k, int64_t total_q,
                           int64_t total_k, int QH, int KH) {
  if (int rc = validate(dtype, E, max_q, max_k, QH, KH, 1)) return rc;
  if (dtype == NNOP_F32)
    return fail(NNOP_ERR_DTYPE, "packed variable-length attention supports Float16 / BFloat16 only");
  if (E != 64 && E != 128)
    return fail(NNOP_ERR_UNSUPPORTED_E,
                "packed variable-length attention supports embedding dims 64 and 128, got `%d`.", E);
  if (nseq < 0 || total_q < 0 || total_k < 0 || total_q > 0x7fffffff || total_k > 0x7fffffff)
    return fail(NNOP_ERR_SHAPE, "Invalid packed shape nseq=%d total_q=%lld total_k=%lld.", nseq,
                static_cast<long long>(total_q), static_cast<long long>(total_k));
  return NNOP_OK;
}

extern "C" int nnop_flash_attn_bwd(void* dq, void* dk, void* dv, void* dpair, const void* dO,
                                   const void* o, const float* lse, const void* q, const void* k,
                                   const void* v, const void* pair, const uint8_t* kpad_mask,
                                   int dtype, int E, int QL, int KL, int QH, int KH, int B,
                                   int causal, float scale, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  return flash_attn_bwd_impl(dq, dk, dv, dpair, dO, o, lse, q, k, v, pair, kpad_mask, dtype, E, QL, KL, QH,
                             KH, B, causal, scale, workspace, workspace_bytes, stream, nullptr);
}

extern "C" int nnop_flash_attn_bwd_reuse_pair(void* dq, void* dk, void* dv, void* dpair, const void* dO,
                                              const void* o, const float* lse, const void* q,
                                              const void* k, const void* v, const void* pair,
                                              const uint8_t* kpad_mask, int dtype, int E, int QL, int KL,
                                              int QH, int KH, int B, int causal, float scale,
                                              void* workspace, size_t workspace_bytes, void* stream,
                                              const void* pair_head_major) {
  return flash_attn_bwd_impl(dq, dk, dv, dpair, dO, o, lse, q, k, v, pair, kpad_mask, dtype, E, QL, KL, QH,
                             KH, B, causal, scale, workspace, workspace_bytes, stream, pair_head_major);
}

extern "C" int nnop_flash_attn_varlen_fwd_ws(void* o, float* lse, const void* q, const void* k,
                                          const void* v, const int32_t* cu_seqlens_q,
                                          const int32_t* cu_seqlens_k, int nseq, int max_seqlen_q,
                                          int max_seqlen_k, int64_t total_q, int64_t total_k,
                                          int dtype, int E, int QH, int KH, int causal, float scale,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  clear_error();
  if (int rc = validate_varlen(dtype, E, nseq, max_seqlen_q, max_seqlen_k, total_q, total_k, QH, KH))
    return rc;
  if (nseq == 0 || total_q == 0 || max_seqlen_q == 0) return NNOP_OK;
  if (!o || !lse || !q || !cu_seqlens_q || !cu_seqlens_k || (total_k > 0 && (!k || !v)))
    return fail(NNOP_ERR_ARG, "NULL pointer");
  AttnParams p{};
  p.o = o; p.lse = lse; p.q = q; p.k = k; p.v = v;
  p.dtype = dtype; p.E = E; p.QL = max_seqlen_q; p.KL = max_seqlen_k; p.QH = QH; p.KH = KH; p.B = 1;
  p.causal = causal ? 1 : 0; p.scale = scale;
  p.stream = static_cast<cudaStream_t>(stream);
  p.cu_q = cu_seqlens_q; p.cu_k = cu_seqlens_k; p.nseq = nseq; p.total_q = total_q;
  p.total_k = total_k > 0 ? total_k : 1;  // a TMA map needs a non-empty extent; no key is ever read
  if (workspace && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && workspace_bytes >= kFwdCounterBytes)
    p.fwd_ws = workspace;   // tile counter: lets batches of short sequences take the persistent forward
  if (!attn_sm100_supported(p, false))
    return fail(NNOP_ERR_ARG, "packed attention: pointers must be 16-byte aligned, heads <= 65535");
  g_last_path = 1;
  return attn_sm100_fwd(p);
}

extern "C" size_t nnop_flash_attn_varlen_fwd_workspace_bytes(int dtype, int E, int nseq, int64_t total_q, int QH) {
  (void)dtype;
  if (E <= 0 || nseq <= 0 || total_q <= 0 || QH <= 0) return 0;
  return kFwdCounterBytes;
}

extern "C" int nnop_flash_attn_varlen_fwd(void* o, float* lse, const void* q, const void* k, const void* v,
                                          const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int nseq,
                                          int max_seqlen_q, int max_seqlen_k, int64_t total_q, int64_t total_k,
                                          int dtype, int E, int QH, int KH, int causal, float scale, void* stream) {
  return nnop_flash_attn_varlen_fwd_ws(o, lse, q, k, v, cu_seqlens_q, cu_seqlens_k, nseq, max_seqlen_q, max_seqlen_k,
                                       total_q, total_k, dtype, E, QH, KH, causal, scale, nullptr, 0, stream);
}

extern "C" size_t nnop_flash_attn_varlen_bwd_workspace_bytes(int dtype, int E, int nseq,
                                                             int64_t total_q, int QH) {
  (void)dtype;
  if (E <= 0 || nseq <= 0 || total_q <= 0 || QH <= 0) return 0;
  return attn_sm100_bwd_packed_workspace_bytes(E, total_q, nseq, QH);
}

extern "C" int nnop_flash_attn_varlen_bwd(void* dq, void* dk, void* dv, const void* dO,
                                          const void* o, const float* lse, const void* q,
                                          const void* k, const void* v,
                                          const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k,
                                          int nseq, int max_seqlen_q, int max_seqlen_k,
                                          int64_t total_q, int64_t total_k, int dtype, int E, int QH,
                                          int KH, int causal, float scale, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  clear_error();
  if (int rc = validate_varlen(dtype, E, nseq, max_seqlen_q, max_seqlen_k, total_q, total_k, QH, KH))
    return rc;
  if (nseq == 0 || (total_q == 0 && total_k == 0)) return NNOP_OK;
  if (!cu_seqlens_q || !cu_seqlens_k || (total_q > 0 && (!dq || !dO || !o || !lse || !q)) ||
      (total_k > 0 && (!dk || !dv || !k || !v)))
    return fail(NNOP_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (total_q == 0 || max_seqlen_q == 0 || total_k == 0 || max_seqlen_k == 0) {  // all gradients are zero
    if (total_k > 0) {
      const size_t kvb = static_cast<size_t>(KH) * total_k * E * dtype_size(dtype);
      NNOP_CUDA_CHECK(cudaMemsetAsync(dk, 0, kvb, st));
      NNOP_CUDA_CHECK(cudaMemsetAsync(dv, 0, kvb, st));
    }
    if (total_q > 0)
      NNOP_CUDA_CHECK(cudaMemsetAsync(dq, 0, static_cast<size_t>(QH) * total_q * E * dtype_size(dtype), st));
    return NNOP_OK;
  }
  const size_t need = nnop_flash_attn_varlen_bwd_workspace_bytes(dtype, E, nseq, total_q, QH);
  if (!workspace || workspace_bytes < need)
    return fail(NNOP_ERR_WORKSPACE, "packed flash attention backward needs a %zu-byte workspace, got %zu",
                need, workspace_bytes);
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0)
    return fail(NNOP_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  AttnParams p{};
  p.o = const_cast<void*>(o); p.lse = const_cast<float*>(lse);
  p.q = q; p.k = k; p.v = v; p.dq = dq; p.dk = dk; p.dv = dv; p.dO = dO;
  p.delta = static_cast<float*>(workspace);
  p.dtype = dtype; p.E = E; p.QL = max_seqlen_q; p.KL = max_seqlen_k; p.QH = QH; p.KH = KH; p.B = 1;
  p.causal = causal ? 1 : 0; p.scale = scale; p.stream = st;
  p.cu_q = cu_seqlens_q; p.cu_k = cu_seqlens_k; p.nseq = nseq; p.total_q = total_q; p.total_k = total_k;
  if (!attn_sm100_supported(p, true))
    return fail(NNOP_ERR_ARG, "packed attention: pointers must be 16-byte aligned, heads <= 65535");
  g_last_path = 1;
  return attn_sm100_bwd(p);
}
