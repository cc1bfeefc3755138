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
// Tiles that lie fully inside the arrays (almost all of them) run without per-element predicates and
// with the row bases hoisted: the generic form spent two thirds of its issue slots on index math.
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
  T* dst_b = out + static_cast<int64_t>(b) * QH * QL * KLp;
  const bool interior = k0 + kTP <= KL && c0 + kTP <= C;
  if (interior) {
    const T* sp = src + static_cast<int64_t>(k0 + threadIdx.y) * C + c0 + threadIdx.x;
#pragma unroll
    for (int i = 0; i < kTP / 8; ++i) {
      tile[threadIdx.y + 8 * i][threadIdx.x] = sp[0];
      tile[threadIdx.y + 8 * i][threadIdx.x + 32] = sp[32];
      sp += static_cast<int64_t>(8) * C;
    }
    __syncthreads();
    int c = c0 + threadIdx.y;
    int q = c / QH, h = c - q * QH;          // one division per thread; then c advances by 8
    const int dq8 = 8 / QH, dh8 = 8 - dq8 * QH;
#pragma unroll
    for (int i = 0; i < kTP / 8; ++i) {
      T* dp = dst_b + (static_cast<int64_t>(h) * QL + q) * KLp + k0 + threadIdx.x;
      dp[0] = tile[threadIdx.x][threadIdx.y + 8 * i];
      dp[32] = tile[threadIdx.x + 32][threadIdx.y + 8 * i];
      q += dq8; h += dh8;
      if (h >= QH) { h -= QH; ++q; }
    }
    return;
  }
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
      T* dst = dst_b + (static_cast<int64_t>(h) * QL + q) * KLp;
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
  // tiles fully inside the arrays and fully ABOVE the diagonal (half of a causal problem): nothing to read,
  // no index math -- plain zero stores with a hoisted base (they used to take the generic path below, which
  // spends ~80 % of its issue slots on per-element divisions and predicates)
  if (all_dead && k0 + kTP <= KL && c0 + kTP <= C) {
    T* dp = out + static_cast<int64_t>(b) * KL * C + static_cast<int64_t>(k0 + threadIdx.y) * C + c0 + threadIdx.x;
#pragma unroll
    for (int i = 0; i < kTP / 8; ++i) {
      dp[0] = T(0);
      dp[32] = T(0);
      dp += static_cast<int64_t>(8) * C;
    }
    return;
  }
  // tiles fully inside the arrays and fully below the diagonal: no per-element predicates, hoisted bases
  if (k0 + kTP <= KL && c0 + kTP <= C && (!causal || k0 + kTP - 1 <= c0 / QH)) {
    int c = c0 + threadIdx.y;
    int q = c / QH, h = c - q * QH;
    const int dq8 = 8 / QH, dh8 = 8 - dq8 * QH;
    const T* src_b = in + static_cast<int64_t>(b) * QH * QL * KLp;
#pragma unroll
    for (int i = 0; i < kTP / 8; ++i) {
      const T* sp = src_b + (static_cast<int64_t>(h) * QL + q) * KLp + k0 + threadIdx.x;
      tile[threadIdx.y + 8 * i][threadIdx.x] = keep[0] ? sp[0] : T(0);
      tile[threadIdx.y + 8 * i][threadIdx.x + 32] = keep[1] ? sp[32] : T(0);
      q += dq8; h += dh8;
      if (h >= QH) { h -= QH; ++q; }
    }
    __syncthreads();
    T* dp = out + static_cast<int64_t>(b) * KL * C + static_cast<int64_t>(k0 + threadIdx.y) * C + c0 + threadIdx.x;
#pragma unroll
    for (int i = 0; i < kTP / 8; ++i) {
      dp[0] = tile[threadIdx.x][threadIdx.y + 8 * i];
      dp[32] = tile[threadIdx.x + 32][threadIdx.y + 8 * i];
      dp += static_cast<int64_t>(8) * C;
    }
    return;
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

// 16-bit payloads, C even: the same tile transposes moving two elements per access (a 2-byte access
// per lane only reaches half the bandwidth).  Pairs run along c on the (B, KL, C) side and along k on
// the head-major side; rows of both sides are then 4-byte aligned (C even, KLp a multiple of 32).
__global__ void __launch_bounds__(256)
pair_to_head_major_16x2_kernel(uint16_t* __restrict__ out, const uint16_t* __restrict__ in, int KL, int QL,
                               int QH, int KLp, int causal) {
  __shared__ __align__(4) uint16_t tile[kTP][kTP + 2];
  const int b = blockIdx.z;
  const int C = QL * QH;
  const int k0 = blockIdx.x * kTP, c0 = blockIdx.y * kTP;
  if (causal && k0 >= (c0 + kTP - 1) / QH + 256) return;
  const uint16_t* src = in + static_cast<int64_t>(b) * KL * C;
  const int c = c0 + 2 * threadIdx.x;  // this lane's pair of columns
  if (k0 + kTP <= KL && c0 + kTP <= C) {   // interior tile: no predicates, hoisted bases
    const uint16_t* sp = src + static_cast<int64_t>(k0 + threadIdx.y) * C + c;
#pragma unroll
    for (int i = 0; i < kTP / 8; ++i) {
      *reinterpret_cast<uint32_t*>(&tile[threadIdx.y + 8 * i][2 * threadIdx.x]) = *reinterpret_cast<const uint32_t*>(sp);
      sp += static_cast<int64_t>(8) * C;
    }
    __syncthreads();
    int cc = c0 + threadIdx.y;
    int q = cc / QH, h = cc - q * QH;
    const int dq8 = 8 / QH, dh8 = 8 - dq8 * QH;
    uint16_t* dst_b = out + static_cast<int64_t>(b) * QH * QL * KLp + k0 + 2 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < kTP / 8; ++i) {
      const int r = threadIdx.y + 8 * i;
      const uint32_t v = static_cast<uint32_t>(tile[2 * threadIdx.x][r]) |
                         (static_cast<uint32_t>(tile[2 * threadIdx.x + 1][r]) << 16);
      *reinterpret_cast<uint32_t*>(dst_b + (static_cast<int64_t>(h) * QL + q) * KLp) = v;
      q += dq8; h += dh8;
      if (h >= QH) { h -= QH; ++q; }
    }
    return;
  }
#pragma unroll
  for (int r = threadIdx.y; r < kTP; r += 8) {
    const int k = k0 + r;
    uint32_t v = 0u;
    if (k < KL && c < C) v = *reinterpret_cast<const uint32_t*>(src + static_cast<int64_t>(k) * C + c);
    *reinterpret_cast<uint32_t*>(&tile[r][2 * threadIdx.x]) = v;
  }
  __syncthreads();
  const int k = k0 + 2 * threadIdx.x;  // this lane's pair of keys
#pragma unroll
  for (int r = threadIdx.y; r < kTP; r += 8) {
    const int cc = c0 + r;
    if (cc < C && k < KLp) {
      const int q = cc / QH, h = cc % QH;
      const uint32_t v = static_cast<uint32_t>(tile[2 * threadIdx.x][r]) |
                         (static_cast<uint32_t>(tile[2 * threadIdx.x + 1][r]) << 16);
      *reinterpret_cast<uint32_t*>(out + ((static_cast<int64_t>(b) * QH + h) * QL + q) * KLp + k) = v;
    }
  }
}

__global__ void __launch_bounds__(256)
dpair_from_head_major_16x2_kernel(uint16_t* __restrict__ out, const uint16_t* __restrict__ in,
                                  const uint8_t* __restrict__ kpad, int KL, int QL, int QH, int KLp, int causal) {
  __shared__ __align__(4) uint16_t tile[kTP][kTP + 2];
  const int b = blockIdx.z;
  const int C = QL * QH;
  const int k0 = blockIdx.x * kTP, c0 = blockIdx.y * kTP;
  const bool all_dead = causal && k0 > (c0 + kTP - 1) / QH;
  const int k = k0 + 2 * threadIdx.x;
  bool keep0 = k < KL, keep1 = k + 1 < KL;
  if (kpad) {
    keep0 = keep0 && kpad[static_cast<int64_t>(b) * KL + k] != 0;
    keep1 = keep1 && kpad[static_cast<int64_t>(b) * KL + k + 1] != 0;
  }
  if (all_dead && k0 + kTP <= KL && c0 + kTP <= C) {   // interior tile fully above the diagonal: zeros, no reads
    uint16_t* dp = out + static_cast<int64_t>(b) * KL * C + static_cast<int64_t>(k0 + threadIdx.y) * C + c0 +
                   2 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < kTP / 8; ++i) {
      *reinterpret_cast<uint32_t*>(dp) = 0u;
      dp += static_cast<int64_t>(8) * C;
    }
    return;
  }
  if (k0 + kTP <= KL && c0 + kTP <= C && (!causal || k0 + kTP - 1 <= c0 / QH)) {
    // interior tile fully below the diagonal: only the key padding mask can kill an entry
    const uint32_t keepm = (keep0 ? 0x0000ffffu : 0u) | (keep1 ? 0xffff0000u : 0u);
    int cc = c0 + threadIdx.y;
    int q = cc / QH, h = cc - q * QH;
    const int dq8 = 8 / QH, dh8 = 8 - dq8 * QH;
    const uint16_t* src_b = in + static_cast<int64_t>(b) * QH * QL * KLp + k;
#pragma unroll
    for (int i = 0; i < kTP / 8; ++i) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(src_b + (static_cast<int64_t>(h) * QL + q) * KLp);
      *reinterpret_cast<uint32_t*>(&tile[threadIdx.y + 8 * i][2 * threadIdx.x]) = v & keepm;
      q += dq8; h += dh8;
      if (h >= QH) { h -= QH; ++q; }
    }
    __syncthreads();
    uint16_t* dp = out + static_cast<int64_t>(b) * KL * C + static_cast<int64_t>(k0 + threadIdx.y) * C + c0 +
                   2 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < kTP / 8; ++i) {
      const int r = threadIdx.y + 8 * i;
      const uint32_t v = static_cast<uint32_t>(tile[2 * threadIdx.x][r]) |
                         (static_cast<uint32_t>(tile[2 * threadIdx.x + 1][r]) << 16);
      *reinterpret_cast<uint32_t*>(dp) = v;
      dp += static_cast<int64_t>(8) * C;
    }
    return;
  }
#pragma unroll
  for (int r = threadIdx.y; r < kTP; r += 8) {
    const int cc = c0 + r;
    const int q = cc / QH, h = cc % QH;
    uint32_t v = 0u;
    if (!all_dead && cc < C && k < KLp)
      v = *reinterpret_cast<const uint32_t*>(in + ((static_cast<int64_t>(b) * QH + h) * QL + q) * KLp + k);
    if (!keep0 || (causal && k > q)) v &= 0xffff0000u;
    if (!keep1 || (causal && k + 1 > q)) v &= 0x0000ffffu;
    *reinterpret_cast<uint32_t*>(&tile[r][2 * threadIdx.x]) = v;   // tile[c][k]
  }
  __syncthreads();
  uint16_t* dst = out + static_cast<int64_t>(b) * KL * C;
  const int c = c0 + 2 * threadIdx.x;
#pragma unroll
  for (int r = threadIdx.y; r < kTP; r += 8) {
    const int kk = k0 + r;
    if (kk < KL && c < C) {
      const uint32_t v = static_cast<uint32_t>(tile[2 * threadIdx.x][r]) |
                         (static_cast<uint32_t>(tile[2 * threadIdx.x + 1][r]) << 16);
      *reinterpret_cast<uint32_t*>(dst + static_cast<int64_t>(kk) * C + c) = v;
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
  else if (C % 2 == 0 && (reinterpret_cast<uintptr_t>(a.pair) & 3) == 0)  // 16-bit payloads are moved bit for bit
    pair_to_head_major_16x2_kernel<<<grid, block, 0, a.stream>>>(
        static_cast<uint16_t*>(a.pair_t), static_cast<const uint16_t*>(a.pair), a.KL, a.QL, a.QH, a.KLp,
        a.causal);
  else
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
  else if (C % 2 == 0 && (reinterpret_cast<uintptr_t>(a.dpair) & 3) == 0)
    dpair_from_head_major_16x2_kernel<<<grid, block, 0, a.stream>>>(
        static_cast<uint16_t*>(a.dpair), static_cast<const uint16_t*>(a.dpair_t), a.kpad, a.KL, a.QL, a.QH,
        a.KLp, a.causal);
  else
    dpair_from_head_major_kernel<uint16_t><<<grid, block, 0, a.stream>>>(
        static_cast<uint16_t*>(a.dpair), static_cast<const uint16_t*>(a.dpair_t), a.kpad, a.KL, a.QL, a.QH,
        a.KLp, a.causal);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

}  // namespace nnop
