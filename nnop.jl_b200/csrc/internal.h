// internal.h -- host-side glue shared by the translation units of libnnop_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/nnop_b200.h"
#include "../../include/nnop_b200_diag.h"

namespace nnop {

// thread-local error message; returns `code` so call sites can `return fail(...)`.
int fail(int code, const char* fmt, ...);
void clear_error();

#define NNOP_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      (void)cudaGetLastError(); /* reported here: do not leave it for the next launch check */ \
      return ::nnop::fail(NNOP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                     \
    }                                                                                      \
  } while (0)

#define NNOP_LAUNCH_CHECK()                                                                \
  do {                                                                                     \
    cudaError_t _e = cudaGetLastError();                                                   \
    if (_e != cudaSuccess)                                                                 \
      return ::nnop::fail(NNOP_ERR_CUDA, "kernel launch failed: %s (%s:%d)",               \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                     \
  } while (0)

inline size_t dtype_size(int dtype) { return dtype == NNOP_F32 ? 4 : 2; }
inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

struct AttnParams {
  void* o;
  float* lse;
  const void* q;
  const void* k;
  const void* v;
  const void* pair;
  const uint8_t* kpad;
  // backward only
  void* dq;
  void* dk;
  void* dv;
  void* dpair;
  const void* dO;
  float* delta;     // workspace (B, QH, QL) fp32
  float* dq_accum;  // workspace (B, QH, QL, E) fp32 (tcgen05 path)
  int dtype, E, QL, KL, QH, KH, B, causal;
  float scale;
  cudaStream_t stream;
  // packed variable-length mode (tcgen05 path only): cu_q / cu_k are device arrays of nseq+1
  // int32 row offsets into (QH, total_q, E) / (KH, total_k, E) tensors; QL / KL = max lengths
  const int* cu_q;
  const int* cu_k;
  int nseq;
  int64_t total_q, total_k;
  // forward workspace (nnop_flash_attn_fwd_ws): [hi | lo] fp16 copies of q, k, v for the Float32 path; 16-bit:
  // the persistent forward's tile counter (nullptr => one CTA per q tile)
  void* fwd_ws;
  // pair bias on the tcgen05 path: head-major copy of pair (B, QH, QL, KLp) and, backward, the staging
  // area dpair is produced in (same layout); KLp = KL rounded up to 32 elements
  void* pair_t;
  void* dpair_t;
  int KLp;
  int pair_t_ready;  // backward: pair_t already holds the copy (made by the forward)
};

// one-shot timing hook (api.cu); which: 0 forward kernel, 1 backward main kernel
void timing_begin(int which, cudaStream_t st);
void timing_end(int which, cudaStream_t st);

// attn_generic.cu -- SIMT path: any dtype, any power-of-two E <= 256, pair, kpad, ragged
int attn_generic_fwd(const AttnParams& p);
int attn_generic_bwd(const AttnParams& p);
// delta[b,h,q] = sum_e dO*O  (shared by both backward paths)
int attn_bwd_preprocess(const AttnParams& p);

// attn_fwd_sm100.cu / attn_bwd_sm100.cu -- tcgen05 + TMA path (bf16/f16, E in {64,128})
bool attn_sm100_supported(const AttnParams& p, bool backward);
int attn_sm100_fwd(const AttnParams& p);
void attn_sm100_set_fwd_mode(int mode);  // 0 auto, 1 one CTA per q tile, 2 persistent, 100+n persistent on n CTAs
size_t attn_sm100_fwd_workspace_bytes(int dtype, int E, int QL, int KL, int QH, int KH, int B);
// Float32 tensor-core path (E = 64): every operand tensor is carried as two fp16 terms of x * 2^-e, e the
// tensor's own binary exponent, so that fp16's range never clips or flushes it.  The 256-byte scale block at
// the end of the Float32 workspaces holds |x|max bit patterns [u32 0..3: q, k, v, dO], the exponents
// [i32 4..7], the multipliers the kernels apply to undo the scaling [f32 8 + F32Mult::k*], the input scales
// 2^-e_x [f32 16..19] and the block counter of the |x|max pass [u32 20].
constexpr size_t kF32ScaleBytes = 256;
constexpr size_t kFwdCounterBytes = 256;   // 16-bit forward workspace: tile counter of the persistent kernel
struct F32Mult {
  enum : int {
    kLogits = 0,    // 2^(e_q + e_k): S = kLogits * q' k'^T           (on top of scale * log2e)
    kO = 1,         // 2^e_v: O = kO * P v'
    kDeltaInv = 2,  // 2^-(e_dO + e_v): delta' = delta * kDeltaInv, so that dS' = P o (dP' - delta') is O(1)
    kDV = 3,        // 2^e_dO: dV = kDV * P^T dO'
    kDQ = 4,        // 2^(e_dO + e_v + e_k): dQ = scale * kDQ * dS' k'
    kDK = 5,        // 2^(e_dO + e_v + e_q): dK = scale * kDK * dS'^T q'
    kDPair = 6      // 2^(e_dO + e_v): dpair = kDPair * dS'
  };
};
// 2^-e_x of tensor `which` (0 q, 1 k, 2 v, 3 dO): what the split kernel multiplies by
inline const float* f32_in_scale(const void* block, int which) { return static_cast<const float*>(block) + 16 + which; }
inline const float* f32_mults(const void* block) { return static_cast<const float*>(block) + 8; }
// memset + |x|max of q, k, v (and dO, may be NULL) + exponents / multipliers; n* = element counts
int attn_f32_scales(void* block, const void* q, int64_t nq, const void* k, int64_t nk, const void* v,
                    int64_t nv, const void* dO, int64_t ndo, cudaStream_t st);
// (rows, E) fp32, E in {16, 32, 64} -> (rows, 128) fp16 rows [hi(64) | lo(64)] with x * (*scale_slot) ~ hi + lo
// (columns >= E of each half are written as zeros)
int attn_split_f32_rows(void* out_bf16x2, const void* in_f32, int64_t rows, int E, const float* scale_slot,
                        cudaStream_t st);
// The same for up to four tensors, plus up to two fp32 regions to zero (the backward's dk / dv accumulators), in ONE
// launch: at the reference's README shape the Float32 calls are a dozen 5-10 us nodes around a 0.2-0.7 ms kernel.
struct F32StageJobs {
  void* out[4] = {nullptr, nullptr, nullptr, nullptr};          // (rows, 2 * half width) fp16
  const void* in[4] = {nullptr, nullptr, nullptr, nullptr};     // (rows, E) fp32
  int64_t rows[4] = {0, 0, 0, 0};
  const float* scale[4] = {nullptr, nullptr, nullptr, nullptr};
  void* zero[2] = {nullptr, nullptr};                           // 16-byte aligned fp32 regions ...
  int64_t zero_floats[2] = {0, 0};                              // ... of this many floats (multiples of 4)
};
int attn_stage_f32(const F32StageJobs& jobs, int E, cudaStream_t st);
// attn_bwd_f32_sm100.cu -- Float32 (E = 64) backward on the tensor cores (split-bf16 operands)
size_t attn_f32_bwd_workspace_bytes(int QL, int KL, int QH, int KH, int B);
int attn_f32_bwd(const AttnParams& p);
int attn_sm100_bwd(const AttnParams& p);
bool attn_sm100_bwd_available();
void attn_sm100_set_bwd_pair_mode(int mode);  // 0 auto, 1 CTA pairs, 2 one CTA per tile, 3 persistent, 100+n persistent on n CTAs
size_t attn_sm100_bwd_workspace_bytes(int E, int QL, int QH, int B);
size_t attn_sm100_bwd_packed_workspace_bytes(int E, int64_t total_q, int nseq, int QH);

// attn_pair.cu -- layout changes of the additive bias for the tcgen05 path
inline int pair_klp(int KL) { return (KL + 31) & ~31; }
size_t attn_pair_workspace_bytes(int dtype, int QL, int KL, int QH, int B, bool backward);
int attn_pair_to_head_major(const AttnParams& p);    // pair (B,KL,QL,QH) -> pair_t (B,QH,QL,KLp)
int attn_dpair_from_head_major(const AttnParams& p);  // dpair_t -> dpair, zero where masked

// TMA descriptor helper (api.cu): 3-D map over a row-major (outer, rows, inner) tensor of
// nnop_dtype_t elements, box (box_inner, box_rows, 1), 128-byte swizzle, zero OOB fill.
int make_tmap_3d(void* tmap_out, const void* base, int dtype, uint64_t inner, uint64_t rows,
                 uint64_t outer, uint32_t box_inner, uint32_t box_rows);

}  // namespace nnop
