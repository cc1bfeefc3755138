/*
 * nnop_b200.h -- C ABI of libnnop_b200.so, the B200 (sm_100a) implementation of the
 * NNop.jl hot path.  Every entry point is what a Julia `ccall` in the NNop shim binds in
 * place of a KernelAbstractions launch of the reference (file:line cited per function,
 * relative to pxl-th/NNop.jl v0.2.0).
 *
 * Conventions
 *  - All pointers are DEVICE pointers (CuArray / torch.cuda storage) unless named host_*.
 *    The library never allocates, frees or retains user-visible memory; outputs,
 *    residuals and workspaces are supplied by the caller.
 *  - Arrays are in the reference's column-major layout.  A Julia (E, L, H, B) array is the
 *    same bytes as a row-major (B, H, L, E) array; the comments below give the Julia shape.
 *  - `stream` is a cudaStream_t / CUstream passed as void*.  Calls only enqueue work; they
 *    never synchronise the device.
 *  - Return value: 0 (NNOP_OK) on success, otherwise an nnop_status_t; a human-readable
 *    message for the calling thread is available from nnop_last_error_string().
 *  - dtype: element type T of q/k/v/x...; softmax statistics, lse, rstd, mean are float.
 *  - State: none.  The library keeps no device memory and no mutable process state behind these entry
 *    points (scratch, including the tile counters of the persistent kernels, lives in the caller's
 *    workspace); every call may run concurrently with any other on any stream.  The process-wide kernel
 *    selection switches and the measurement hooks used by tests and A/B timing are NOT part of this ABI:
 *    they are declared in nnop_b200_diag.h.
 */
#ifndef NNOP_B200_H
#define NNOP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NNOP_B200_VERSION 100 /* 0.1.0 */

typedef enum {
  NNOP_F32 = 0,  /* Float32  */
  NNOP_F16 = 1,  /* Float16  */
  NNOP_BF16 = 2  /* BFloat16 */
} nnop_dtype_t;

typedef enum {
  NNOP_OK = 0,
  NNOP_ERR_SHAPE = 1,         /* shape violation; message mirrors src/attention.jl:141-144 */
  NNOP_ERR_DTYPE = 2,
  NNOP_ERR_UNSUPPORTED_E = 3, /* embedding dim not a power of two / out of range */
  NNOP_ERR_WORKSPACE = 4,     /* workspace NULL or too small */
  NNOP_ERR_CUDA = 5,          /* a CUDA runtime / driver call failed */
  NNOP_ERR_ARG = 6            /* NULL pointer or otherwise invalid argument */
} nnop_status_t;

typedef struct {
  int sm_count;
  int cc_major, cc_minor;
  size_t shared_mem_per_block_optin; /* replaces NNop._shared_memory, ext/NNopCUDAExt.jl:6-9 */
  size_t l2_bytes;
  size_t hbm_bytes;
} nnop_device_info_t;

int nnop_version(void);
const char* nnop_last_error_string(void);
/* Replaces NNop.shared_memory / _shared_memory (src/NNop.jl:27-30, ext/NNopCUDAExt.jl:6-9). */
int nnop_device_info(int device, nnop_device_info_t* out);

/* ---------------------------------------------------------------------------------------
 * flash attention forward.  Replaces `_flash_attention` + kernel `_flash_attention_fwd!`
 * (src/attention.jl:133-177, :1-131).
 *   o    (E, QL, QH, B)  T      out
 *   lse  (QL, QH, B)     float  out: m + log(l), natural log, of the scaled+biased+masked
 *                               logits.  Replaces the reference residual pair (ms, ls)
 *                               (src/attention.jl:166-168).  -inf for a fully masked row,
 *                               whose output row is 0 (the reference yields NaN there).
 *   q    (E, QL, QH, B)  T ;  k, v (E, KL, KH, B) T ; QH % KH == 0 (GQA; q-head j reads
 *                               kv-head j / (QH/KH), src/attention.jl:28)
 *   pair       (QH, QL, KL, B) T  or NULL : additive logit bias   (src/attention.jl:59-64)
 *   kpad_mask  (KL, B) uint8 0/1  or NULL : 1 = attend            (src/attention.jl:73-79)
 *   causal     keep k_idx <= q_idx (top-left aligned)             (src/attention.jl:67-72)
 *   scale      logit scale; the reference always passes 1/sqrt(E) (src/attention.jl:154)
 */
int nnop_flash_attn_fwd(void* o, float* lse, const void* q, const void* k, const void* v,
                        const void* pair, const uint8_t* kpad_mask, int dtype, int E, int QL,
                        int KL, int QH, int KH, int B, int causal, float scale, void* stream);

/* Forward with an optional caller-owned workspace.  With a workspace of at least
 * nnop_flash_attn_fwd_workspace_bytes(...) bytes (256-byte aligned), 16-bit problems whose tile queue is deep
 * and short-tiled enough may run the persistent forward kernel (its tile counter lives in the workspace;
 * O and lse are bit-identical either way), and Float32 problems with E in {16, 32, 64} (E = 128 without `pair`)
 * run on the tensor cores: q, k, v are scaled by a power of two and split into two fp16 terms each (x ~ hi + lo,
 * 22 significant bits), S = Qh Kh^T + Qh Kl^T + Ql Kh^T and O = (Ph + Pl)(Vh + Vl) accumulate in fp32;
 * max abs error vs an fp64 evaluation stays below 1e-4.  Without it (or for other shapes) the call
 * is identical to nnop_flash_attn_fwd.  The size query returns 0 where no workspace is used.
 *
 * `pair` (the additive bias, (QH,QL,KL,B) column-major, src/attention.jl:55-62) runs on the tensor
 * cores when the workspace is additionally extended by nnop_flash_attn_pair_workspace_bytes(...,
 * backward = 0) bytes: total = round_up(fwd_workspace_bytes, 256) + pair_workspace_bytes.  The
 * library then keeps a head-major copy of pair there (its layout has the head as the fastest axis,
 * which no tile load can fetch).  Same rule for nnop_flash_attn_bwd with backward = 1 (copy of pair
 * plus the staging area dpair is produced in): total = round_up(bwd_workspace_bytes, 256) +
 * pair_workspace_bytes(..., 1).  With a smaller workspace `pair` is served by the SIMT kernels.
 * (16-bit E = 256 runs its forward on the tensor cores without any workspace; its backward, like that of Float32
 * E = 128, and every `pair` call at those widths are served by the SIMT kernels.) */
size_t nnop_flash_attn_pair_workspace_bytes(int dtype, int QL, int KL, int QH, int B, int backward);
size_t nnop_flash_attn_fwd_workspace_bytes(int dtype, int E, int QL, int KL, int QH, int KH, int B);
int nnop_flash_attn_fwd_ws(void* o, float* lse, const void* q, const void* k, const void* v,
                           const void* pair, const uint8_t* kpad_mask, int dtype, int E, int QL,
                           int KL, int QH, int KH, int B, int causal, float scale, void* workspace,
                           size_t workspace_bytes, void* stream);

/* Workspace needed by nnop_flash_attn_bwd for these dims (delta, fp32 dQ accumulator). */
size_t nnop_flash_attn_bwd_workspace_bytes(int dtype, int E, int QL, int KL, int QH, int KH,
                                           int B);

/* flash attention backward.  Replaces `∇flash_attention` + kernels
 * `_flash_attention_bwd_preprocess!` and `_flash_attention_bwd!`
 * (src/attention_bwd.jl:199-275, :163-197, :1-161).
 *   dq (E,QL,QH,B), dk, dv (E,KL,KH,B)  T out -- fully overwritten (no pre-zeroing needed)
 *   dpair (QH,QL,KL,B) T out or NULL (required iff pair != NULL)
 *   dO, o (E,QL,QH,B) T ; lse from the forward; remaining arguments as in the forward.
 */
int nnop_flash_attn_bwd(void* dq, void* dk, void* dv, void* dpair, const void* dO,
                        const void* o, const float* lse, const void* q, const void* k,
                        const void* v, const void* pair, const uint8_t* kpad_mask, int dtype,
                        int E, int QL, int KL, int QH, int KH, int B, int causal, float scale,
                        void* workspace, size_t workspace_bytes, void* stream);
/* Same, for callers that kept the forward's workspace alive: `pair_head_major` is the head-major copy
 * of `pair` that nnop_flash_attn_fwd_ws left at offset round_up(nnop_flash_attn_fwd_workspace_bytes,
 * 256) of its workspace (same dims, same dtype).  The backward then skips its own layout change of
 * pair and its workspace only needs round_up(bwd_workspace_bytes, 256) + pair_workspace_bytes(..., 0)
 * (the dpair staging area).  NULL = nnop_flash_attn_bwd. */
int nnop_flash_attn_bwd_reuse_pair(void* dq, void* dk, void* dv, void* dpair, const void* dO,
                                   const void* o, const float* lse, const void* q, const void* k,
                                   const void* v, const void* pair, const uint8_t* kpad_mask,
                                   int dtype, int E, int QL, int KL, int QH, int KH, int B,
                                   int causal, float scale, void* workspace, size_t workspace_bytes,
                                   void* stream, const void* pair_head_major);

/* ---------------------------------------------------------------------------------------
 * Packed variable-length flash attention (additive: the reference only has the dense Bool
 * `kpad_mask`, src/attention.jl:73-79, which pays for every padded tile).  nseq sequences are
 * concatenated along L:
 *   q, o, dq, dO (E, total_q, QH) T ;  k, v, dk, dv (E, total_k, KH) T ;  lse (total_q, QH) float
 *   cu_seqlens_q / cu_seqlens_k : nseq+1 int32 DEVICE arrays, cu[0] = 0, cu[nseq] = total;
 *   sequence z owns rows [cu[z], cu[z+1]).  max_seqlen_* >= the longest sequence (sizes the grid).
 *   causal is top-left aligned per sequence (k_idx <= q_idx), as in the dense call.
 * Float16 / BFloat16, E in {64, 128} (tcgen05 path); same math as `_flash_attention` /
 * `∇flash_attention` applied per sequence (src/attention.jl:133-177, src/attention_bwd.jl:199-275).
 */
int nnop_flash_attn_varlen_fwd(void* o, float* lse, const void* q, const void* k, const void* v,
                               const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int nseq,
                               int max_seqlen_q, int max_seqlen_k, int64_t total_q, int64_t total_k,
                               int dtype, int E, int QH, int KH, int causal, float scale,
                               void* stream);
/* Same with a caller-owned workspace (>= nnop_flash_attn_varlen_fwd_workspace_bytes, 256-byte aligned):
 * batches of short sequences then run the persistent forward kernel, whose tile counter lives there. */
size_t nnop_flash_attn_varlen_fwd_workspace_bytes(int dtype, int E, int nseq, int64_t total_q, int QH);
int nnop_flash_attn_varlen_fwd_ws(void* o, float* lse, const void* q, const void* k, const void* v,
                                  const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int nseq,
                                  int max_seqlen_q, int max_seqlen_k, int64_t total_q, int64_t total_k,
                                  int dtype, int E, int QH, int KH, int causal, float scale,
                                  void* workspace, size_t workspace_bytes, void* stream);
size_t nnop_flash_attn_varlen_bwd_workspace_bytes(int dtype, int E, int nseq, int64_t total_q,
                                                  int QH);
int nnop_flash_attn_varlen_bwd(void* dq, void* dk, void* dv, const void* dO, const void* o,
                               const float* lse, const void* q, const void* k, const void* v,
                               const int32_t* cu_seqlens_q, const int32_t* cu_seqlens_k, int nseq,
                               int max_seqlen_q, int max_seqlen_k, int64_t total_q, int64_t total_k,
                               int dtype, int E, int QH, int KH, int causal, float scale,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Helpers of the sequence-sharded ("ring") attention variant (additive: the reference has no
 * multi-GPU path).  A rank attends its local queries to one K/V block per ring step with
 * nnop_flash_attn_fwd and folds the partial result into fp32 accumulators:
 *   nnop_attn_merge: (o_acc, lse_acc) <- logsumexp-weighted combination with (o_part, lse_part);
 *     o_acc (E, rows) float in/out, lse_acc (rows) float in, lse_out (rows) float out (must not
 *     alias lse_acc unless init), o_part (E, rows) T, lse_part (rows) float; init != 0 copies.
 *   nnop_accumulate_f32: acc[n] (+)= float(part[n])  (dq over steps; travelling dk / dv); n a multiple
 *     of the elements in a 128-bit vector of `part` (4 for Float32, 8 for 16-bit).
 *   nnop_store_rows_from_f32: out[:, offset : offset+rows, slab] = T(acc[:, :, slab]) for
 *     acc (E, rows, n_slabs) float and out (E, out_slab_rows, n_slabs) T.
 * E % 8 == 0 and every array pointer 16-byte aligned (128-bit accesses), else NNOP_ERR_SHAPE / _ARG.
 */
int nnop_attn_merge(float* o_acc, float* lse_acc, float* lse_out, const void* o_part,
                    const float* lse_part, int dtype, int E, int64_t rows, int init, void* stream);
int nnop_accumulate_f32(float* acc, const void* part, int dtype, int64_t n, int init, void* stream);
int nnop_store_rows_from_f32(void* out, const float* acc, int dtype, int E, int64_t n_slabs,
                             int64_t rows, int64_t out_slab_rows, int64_t out_row_offset,
                             void* stream);

/* ---------------------------------------------------------------------------------------
 * Sequence-sharded ("ring") flash attention over the GPUs of ONE process (additive: the reference has
 * no multi-GPU path; BASELINE config C5).  ndev ranks; rank r runs on CUDA device devices[r] (a device
 * may appear more than once: ranks then share it, which is how a single-GPU box tests the schedule).
 * Every array argument is a HOST array of ndev DEVICE pointers, one per rank:
 *   q[r], o[r], dq[r], dO[r] (E, Ll, QH, B) T ;  k[r], v[r], dk[r], dv[r] (E, Ll, KH, B) T ;
 *   lse[r] (Ll, QH, B) float -- the log-sum-exp over the WHOLE sequence (forward: out; backward: in,
 *   together with the final o).  Rank r holds rows [r*Ll, (r+1)*Ll) of the sequence; with causal != 0
 *   the zig-zag layout instead: the sequence is cut into 2*ndev chunks of Ll/2 rows and rank r holds
 *   chunks r and 2*ndev-1-r, concatenated (equal work per rank and step, no mask beyond step 0).
 *   workspace[r]: device memory on devices[r], 256-byte aligned, at least *_workspace_bytes(...) each.
 *   streams[r]: cudaStream_t of devices[r] (NULL array or NULL entries = default streams).  Inputs are
 *   read, and outputs complete, in the order of streams[r]; the call only enqueues.
 * K / V blocks are pulled from their owner's tensors over NVLink (cudaMemcpyPeerAsync on copy streams
 * made and released inside the call; peer access is enabled where available) one step ahead of the
 * math; each step is one dense nnop_flash_attn_fwd / _bwd launch sequence; dk / dv partials are pushed
 * to the block's owner and accumulated there in fp32.  Same math as `_flash_attention` /
 * `∇flash_attention` on the gathered sequence (src/attention.jl:133-177, src/attention_bwd.jl:199-275).
 */
size_t nnop_ring_attn_fwd_workspace_bytes(int dtype, int E, int Ll, int QH, int KH, int B, int ndev,
                                          int causal);
int nnop_ring_attn_fwd(void* const* o, float* const* lse, const void* const* q, const void* const* k,
                       const void* const* v, const int* devices, int ndev, int dtype, int E, int Ll,
                       int QH, int KH, int B, int causal, float scale, void* const* workspace,
                       size_t workspace_bytes, void* const* streams);
size_t nnop_ring_attn_bwd_workspace_bytes(int dtype, int E, int Ll, int QH, int KH, int B, int ndev,
                                          int causal);
int nnop_ring_attn_bwd(void* const* dq, void* const* dk, void* const* dv, const void* const* dO,
                       const void* const* o, const float* const* lse, const void* const* q,
                       const void* const* k, const void* const* v, const int* devices, int ndev,
                       int dtype, int E, int Ll, int QH, int KH, int B, int causal, float scale,
                       void* const* workspace, size_t workspace_bytes, void* const* streams);

/* ---------------------------------------------------------------------------------------
 * online softmax over dim 1 of x (N, cols).  Replaces `online_softmax` / `online_softmax!`
 * (src/softmax.jl:60-68, :19-58) and `∇online_softmax` (src/softmax.jl:70-80).
 */
int nnop_softmax_fwd(void* y, const void* x, int dtype, int64_t N, int64_t cols, void* stream);
int nnop_softmax_bwd(void* dx, const void* dy, const void* y, int dtype, int64_t N,
                     int64_t cols, void* stream);

/* ---------------------------------------------------------------------------------------
 * RMS norm over dim 1 of x (emb, n).  Replaces `_rms_norm` / `_rms_norm!`
 * (src/rms_norm.jl:117-137, :3-38) and `∇rms_norm` / `_∇rms_norm!` (:139-169, :43-115).
 *   y (emb,n) T out ; rstd (n) float out (the reference's `rms` residual, :27)
 *   dw_f32 (emb) float out -- Float32 for every T, as the reference (:146)
 */
int nnop_rms_norm_fwd(void* y, float* rstd, const void* x, const void* w, int dtype,
                      int64_t emb, int64_t n, float eps, float offset, void* stream);
size_t nnop_norm_bwd_workspace_bytes(int64_t emb, int64_t n);
int nnop_rms_norm_bwd(void* dx, float* dw_f32, const void* dy, const float* rstd, const void* x,
                      const void* w, int dtype, int64_t emb, int64_t n, float offset,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * layer norm over dim 1 of x (emb, n).  Replaces `_layer_norm` / `_layer_norm!`
 * (src/layer_norm.jl:150-170, :8-63) and `∇layer_norm` / `_∇layer_norm!` (:172-204, :65-148).
 *   mean, rstd (n) float out (reference residuals μ, Σ; Σ holds rstd, :50)
 *   dw, db (emb) T out (eltype(w), eltype(b) in the reference, :179-180)
 */
int nnop_layer_norm_fwd(void* y, float* mean, float* rstd, const void* x, const void* w,
                        const void* b, int dtype, int64_t emb, int64_t n, float eps,
                        void* stream);
int nnop_layer_norm_bwd(void* dx, void* dw, void* db, const void* dy, const float* mean,
                        const float* rstd, const void* x, const void* w, int dtype,
                        int64_t emb, int64_t n, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ---------------------------------------------------------------------------------------
 * Llama RoPE.  Replaces `_llama_rope` / `llama_rope!` (src/rope/llama_rope.jl:69-89, :24-65)
 * out-of-place (fuses the reference's `copy(q)`, `copy(k)` at :75-76).
 *   q_out,q_in (E,L,QH,B) T ; k_out,k_in (E,L,KH,B) T ; cos, sin (E,L,B) float, only rows
 *   1..E/2 are read (:43-44).  sin_sign = +1 forward, -1 backward (:86).
 *   In-place use (q_out == q_in) is allowed.
 */
int nnop_llama_rope(void* q_out, void* k_out, const void* q_in, const void* k_in,
                    const float* cos, const float* sin, int dtype, int E, int64_t L, int QH,
                    int KH, int B, float sin_sign, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NNOP_B200_H */
