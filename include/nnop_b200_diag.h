/*
 * nnop_b200_diag.h -- diagnostics of libnnop_b200.so: kernel-variant switches, the "which path ran" query,
 * the measurement hook of bench.py and a hardware self-test.  Used by tests/, bench.py and the A/B timing
 * scripts.  NOT part of the drop-in ABI (include/nnop_b200.h): the switches are process-wide mutable state
 * (atomics; the timing hook and the last-path query are per calling thread), which the product entry points
 * never need -- a host that only binds nnop_b200.h gets the automatic choices.
 */
#ifndef NNOP_B200_DIAG_H
#define NNOP_B200_DIAG_H

#include "nnop_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Attention kernel selection (diagnostics / tests).  0 = auto (tcgen05 path whenever the
 * problem qualifies), 1 = force the generic SIMT path, 2 = require the tcgen05 path (returns
 * NNOP_ERR_ARG if the problem does not qualify).  Process-wide. */
int nnop_set_attention_path(int mode);
/* 1 if the last flash-attention call on this thread ran the tcgen05 path, else 0. */
int nnop_last_attention_path(void);
/* Backward kernel selection on the tcgen05 path (diagnostics / A-B timing), process-wide; env
 * NNOP_BWD_PAIR gives the initial value.  0 (default): automatic -- dense problems without a key
 * padding mask whose tile queue is at least two rounds deep run the persistent kernel (one CTA per
 * SM, dynamic queue of (kv block, kv head, batch) tiles, epilogue overlapped with the next tile),
 * everything else one CTA per tile; 1: E = 128 dense problems run the experimental CTA-pair kernel
 * (tcgen05 cta_group::2; same results, slower); 2: always one CTA per tile; 3: persistent wherever
 * eligible; 4: persistent CTA pairs that exchange dQ halves over distributed shared memory (same dK / dV,
 * measured 2x slower: DSMEM moves ~20 B/clk per SM); 100+n: persistent on n CTAs (tests).  dK / dV are
 * bit-identical across modes. */
int nnop_set_bwd_pair_mode(int mode);
/* Forward kernel selection on the tcgen05 path (diagnostics / A-B timing), process-wide; env
 * NNOP_FWD_MODE gives the initial value.  0 (default): automatic -- dense 16-bit problems without
 * key padding mask or pair bias whose queue of (256-row q tile, head, batch) tiles is at least two
 * rounds deep and whose tiles are short enough for the per-tile fixed cost to matter (E = 64, or
 * QL <= 2048) run the persistent kernel (one CTA per SM, dynamic tile queue, Q / K / V of the next
 * tile loaded under the current one, O stored through private staging); 1: always one CTA per q
 * tile; 2: persistent wherever eligible; 100+n: persistent on n CTAs (tests).  O and lse are
 * bit-identical across these modes.  3: experiment -- one CTA per q tile with two softmax warps per 32
 * query rows (E = 128, dense layout; slower; O within one ulp of the others).  (The persistent forward needs the workspace of nnop_flash_attn_fwd_ws /
 * nnop_flash_attn_varlen_fwd_ws for its tile counter; without one the call runs one CTA per q tile.) */
int nnop_set_fwd_mode(int mode);

/* ---------------------------------------------------------------------------------------
 * Measurement hook (bench.py): the next tcgen05 attention launch of kind `which` (0 = forward
 * kernel, 1 = backward main kernel) made by the calling thread records `start_event` right
 * before and `stop_event` right after that one kernel, on the launch stream.  Events are
 * cudaEvent_t handles passed as void*; the hook is one-shot (cleared once used); NULLs clear it.
 */
int nnop_set_timing_events(int which, void* start_event, void* stop_event);

/* ---------------------------------------------------------------------------------------
 * Hardware self-test of the tcgen05/TMA building blocks (diagnostics; used by tests).
 * Runs one 128x128x128 bf16 GEMM through an operand form the attention kernels use and
 * writes the 128x128 fp32 result to d_out; a, b are 128x128 bf16 row-major device buffers.
 *   which 0: A B^T, TMA + K-major smem operands        which 1: A B, A in TMEM, B MN-major
 *   which 2: A B^T, thread-written swizzled smem       which 3: A^T B, both MN-major
 *   which 4: A B, A K-major, B MN-major
 */
int nnop_selftest_umma(float* d_out, const void* a, const void* b, int which, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NNOP_B200_DIAG_H */
