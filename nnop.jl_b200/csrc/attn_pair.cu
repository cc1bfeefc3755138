// attn_pair.cu -- the additive attention bias `pair` on the tcgen05 path.
// The reference stores pair / dpair as (QH, QL, KL, B) column-major (src/attention.jl:55-62,
// src/attention_bwd.jl:120-131), i.e. row-major (B, KL, QL, QH) with the head as the fastest
// axis: one head's (query x key) tile is strided by QH elements in both directions, which neither
// TMA boxes nor per-row vector loads can fetch efficiently.  The tensor-core kernels therefore work
// on a head-major copy (B, QH, QL, KLp) (KLp = KL rounded up to 32 so that rows are 16-byte
// aligned for TMA): the forward streams (128 query) x (128 byte) boxes of it, the backward reads
// and writes it with lanes along the key axis.  Both layout changes are HBM-bound tile transposes.
#include "common.cuh"
#include "internal.h"

namespace nnop {
namespace {

constexpr int kTP = 64;  // tile edge: 16 independent loads per thread keep enough bytes in flight

// in (B, KL, C) with C = QL * QH, c = q * QH + h  ->  out (B, QH, QL, KLp)
template <typename T>
__global__ void __launch_bounds__(256)
pair_to_head_major_kernel(T* __restrict__ out, const T* __restrict__ in, int KL, int QL, int QH, int KLp,
                          int causal) {
  __shared__ T tile[kTP][kTP + 1];
  const int b = blockIdx.z;
  const int C = QL * QH;
  const int k0 = blockIdx.x * kTP, c0 = blockIdx.y * kTP;
  // causal: a query row q only ever sees keys below q + 256 (forward: its 256-row tile's diagonal
  // blocks; backward: 128-row blocks), and everything above the diagonal is overwritten by the mask
  if (causal && k0 >= (c0 + kTP - 1) / QH + 256) return;
  const T* src = in + static_cast<int64_t>(b) * KL * C;
#pragma unroll
  for (int r = threadIdx.y; r < kTP; r += 8) {
    const int k = k0 + r;
#pragma unroll
    for (int cc = 0; cc < kTP; cc += 32) {
      const int c = c0 + cc + threadIdx.x;
      tile[r][cc + threadIdx.x] = (k < KL && c < C) ? src[static_cast<int64_t>(k) * C + c] : T(0);
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = threadIdx.y; r < kTP; r += 8) {
    const int c = c0 + r;
    if (c < C) {
      const int q = c / QH, h = c % QH;
      T* dst = out + ((static_cast<int64_t>(b) * QH + h) * QL + q) * KLp;
#pragma unroll
      for (int kk = 0; kk < kTP; kk += 32) {
        const int k = k0 + kk + threadIdx.x;
        if (k < KLp) dst[k] = tile[kk + threadIdx.x][r];
      }
    }
  }
}

// in (B, QH, QL, KLp) -> out (B, KL, C); entries the backward never visits are defined as zero:
// causal k > q (src/attention.jl:70) and padded keys (src/attention.jl:73-79)
template <typename T>
__global__ void __launch_bounds__(256)
dpair_from_head_major_kernel(T* __restrict__ out, const T* __restrict__ in, const uint8_t* __restrict__ kpad,
                             int KL, int QL, int QH, int KLp, int causal) {
  __shared__ T tile[kTP][kTP + 1];
  const int b = blockIdx.z;
  const int C = QL * QH;
  const int k0 = blockIdx.x * kTP, c0 = blockIdx.y * kTP;
  const bool all_dead = causal && k0 > (c0 + kTP - 1) / QH;  // whole tile above the diagonal: zeros
  bool keep[kTP / 32];
#pragma unroll
  for (int kk = 0; kk < kTP; kk += 32) {
    const int k = k0 + kk + threadIdx.x;
    keep[kk / 32] = k < KL && (!kpad || kpad[static_cast<int64_t>(b) * KL + k] != 0);
  }
#pragma unroll
  for (int r = threadIdx.y; r < kTP; r += 8) {
    const int c = c0 + r;
    const int q = c / QH, h = c % QH;
    const T* src = in + ((static_cast<int64_t>(b) * QH + h) * QL + q) * KLp;
#pragma unroll
    for (int kk = 0; kk < kTP; kk += 32) {
      const int k = k0 + kk + threadIdx.x;
      const bool live = !all_dead && c < C && keep[kk / 32] && !(causal && k > q);
      tile[r][kk + threadIdx.x] = live ? src[k] : T(0);
    }
  }
  __syncthreads();
  T* dst = out + static_cast<int64_t>(b) * KL * C;
#pragma unroll
  for (int r = threadIdx.y; r < kTP; r += 8) {
    const int k = k0 + r;
    if (k < KL) {
#pragma unroll
      for (int cc = 0; cc < kTP; cc += 32) {
        const int c = c0 + cc + threadIdx.x;
        if (c < C) dst[static_cast<int64_t>(k) * C + c] = tile[cc + threadIdx.x][r];
      }
    }
  }
}

}  // namespace

size_t attn_pair_workspace_bytes(int dtype, int QL, int KL, int QH, int B, bool backward) {
  if (QL <= 0 || KL <= 0 || QH <= 0 || B <= 0) return 0;
  const size_t one = (static_cast<size_t>(B) * QH * QL * pair_klp(KL) * dtype_size(dtype) + 255) &
                     ~static_cast<size_t>(255);
  return backward ? 2 * one : one;
}

int attn_pair_to_head_major(const AttnParams& a) {
  const int C = a.QL * a.QH;
  dim3 grid((a.KLp + kTP - 1) / kTP, (C + kTP - 1) / kTP, a.B), block(32, 8);
  if (a.dtype == NNOP_F32)
    pair_to_head_major_kernel<float><<<grid, block, 0, a.stream>>>(
        static_cast<float*>(a.pair_t), static_cast<const float*>(a.pair), a.KL, a.QL, a.QH, a.KLp, a.causal);
  else  // 16-bit payloads are moved bit for bit
    pair_to_head_major_kernel<uint16_t><<<grid, block, 0, a.stream>>>(
        static_cast<uint16_t*>(a.pair_t), static_cast<const uint16_t*>(a.pair), a.KL, a.QL, a.QH, a.KLp,
        a.causal);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

int attn_dpair_from_head_major(const AttnParams& a) {
  const int C = a.QL * a.QH;
  dim3 grid((a.KL + kTP - 1) / kTP, (C + kTP - 1) / kTP, a.B), block(32, 8);
  if (a.dtype == NNOP_F32)
    dpair_from_head_major_kernel<float><<<grid, block, 0, a.stream>>>(
        static_cast<float*>(a.dpair), static_cast<const float*>(a.dpair_t), a.kpad, a.KL, a.QL, a.QH, a.KLp,
        a.causal);
  else
    dpair_from_head_major_kernel<uint16_t><<<grid, block, 0, a.stream>>>(
        static_cast<uint16_t*>(a.dpair), static_cast<const uint16_t*>(a.dpair_t), a.kpad, a.KL, a.QL, a.QH,
        a.KLp, a.causal);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

}  // namespace nnop
