// rowwise.cu -- the HBM-bound NNop ops: online_softmax, rms_norm, layer_norm (fwd + bwd).
//
// Data layout: the reference's (emb, n) column-major matrix is n contiguous rows of `emb`
// elements; every op normalises/reduces one such row.  Design (B200): 128-bit coalesced
// loads, the row is read from HBM exactly once and kept in registers between the reduction
// and the normalisation (the reference reads it 2-3x: src/rms_norm.jl:16-36,
// src/layer_norm.jl:21-61, src/softmax.jl:27-56), warp-shuffle reductions with one smem hop
// (replacing @groupreduce, src/groupreduce.jl:13-43), fp32 math for every T.
//   * rows of <= 256 vectors: one warp per row, 8 rows per CTA (no block barrier at all)
//   * rows of <= 2048 vectors (1024 in backward): one 256-thread CTA per row
//   * anything else / unaligned / emb not a multiple of the vector width: scalar fallback
// Backward dw/db: each CTA of a persistent grid owns a fixed set of columns per thread,
// accumulates its rows' contributions in registers and writes one fp32 partial row; a second
// kernel reduces the partial rows (deterministic, no atomics; replaces the reference's
// (cld(n,4), emb) partials + sum(dims=1), src/rms_norm.jl:146,166).
#include "common.cuh"
#include "internal.h"

namespace nnop {
namespace {

constexpr int kThreads = 256;
// resident CTAs per SM the short-row backward kernels are compiled for (3 => 80 registers and a few spilled
// bytes in the layer-norm variant: measured slower, 64 vs 53 us at config C3; 2 => 110 registers, none)
#ifndef NNOP_ROWWISE_BWD_MINB
#define NNOP_ROWWISE_BWD_MINB 2
#endif

template <typename T>
struct VecIO {
  static constexpr int N = 16 / sizeof(T);
};

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float (&out)[VecIO<T>::N]);
template <>
__device__ __forceinline__ void load_vec<float>(const float* p, float (&out)[4]) {
  float4 v = *reinterpret_cast<const float4*>(p);
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <>
__device__ __forceinline__ void load_vec<__half>(const __half* p, float (&out)[8]) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    out[2 * i] = f.x; out[2 * i + 1] = f.y;
  }
}
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16>(const __nv_bfloat16* p, float (&out)[8]) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    out[2 * i] = f.x; out[2 * i + 1] = f.y;
  }
}
// raw 16-byte load (kept packed while a prefetched row waits in registers) + unpack
template <typename T>
__device__ __forceinline__ uint4 load_raw(const T* p) {
  return *reinterpret_cast<const uint4*>(p);
}
template <typename T>
__device__ __forceinline__ void unpack_raw(const uint4& v, float (&out)[VecIO<T>::N]);
template <>
__device__ __forceinline__ void unpack_raw<float>(const uint4& v, float (&out)[4]) {
  out[0] = __uint_as_float(v.x); out[1] = __uint_as_float(v.y);
  out[2] = __uint_as_float(v.z); out[3] = __uint_as_float(v.w);
}
template <>
__device__ __forceinline__ void unpack_raw<__half>(const uint4& v, float (&out)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    out[2 * i] = f.x; out[2 * i + 1] = f.y;
  }
}
template <>
__device__ __forceinline__ void unpack_raw<__nv_bfloat16>(const uint4& v, float (&out)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    out[2 * i] = f.x; out[2 * i + 1] = f.y;
  }
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float (&in)[VecIO<T>::N]);
template <>
__device__ __forceinline__ void store_vec<float>(float* p, const float (&in)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
}
template <>
__device__ __forceinline__ void store_vec<__half>(__half* p, const float (&in)[8]) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(in[2 * i], in[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16>(__nv_bfloat16* p, const float (&in)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}

// sum / max over the TPR threads that share a row (TPR == 32: a warp; TPR == 256: the CTA).
// The block form hops through `red`, a [2][2][8] scratch indexed by a parity the caller flips
// per reduction: consecutive reductions use different halves, so ONE barrier per reduction is
// enough (a thread can only overwrite half p again after passing the barrier of the reduction
// in between, which every thread reaches after it has read half p).  The rows of these kernels
// are a dependent chain load -> reduce -> normalise -> store with few warps per SM, so each
// barrier removed is latency off the chain (r01: two barriers per reduction, four per
// layer-norm row).
using RedBuf = float[2][2][16];   // up to 16 warps per CTA (the 512-thread backward)
template <int TPR, int NW = kThreads / 32>
__device__ __forceinline__ float row_sum(float v, RedBuf& red, int& parity) {
  v = warp_sum(v);
  if constexpr (TPR == 32) return v;
  float* r = red[parity][0];
  parity ^= 1;
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < NW; ++i) t += r[i];
  return t;
}
// two sums in one hop (layer-norm backward: mean(w d xh) and mean(w d))
template <int TPR, int NW = kThreads / 32>
__device__ __forceinline__ void row_sum2(float& a, float& b, RedBuf& red, int& parity) {
  a = warp_sum(a);
  b = warp_sum(b);
  if constexpr (TPR == 32) return;
  float(*r)[16] = red[parity];
  parity ^= 1;
  if ((threadIdx.x & 31) == 0) {
    r[0][threadIdx.x >> 5] = a;
    r[1][threadIdx.x >> 5] = b;
  }
  __syncthreads();
  float ta = 0.f, tb = 0.f;
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    ta += r[0][i];
    tb += r[1][i];
  }
  a = ta;
  b = tb;
}
template <int TPR>
__device__ __forceinline__ float row_max(float v, RedBuf& red, int& parity) {
  v = warp_max(v);
  if constexpr (TPR == 32) return v;
  float* r = red[parity][0];
  parity ^= 1;
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = r[0];
#pragma unroll
  for (int i = 1; i < kThreads / 32; ++i) t = fmaxf(t, r[i]);
  return t;
}

// ---------------------------------------------------------------------------------------
// packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2 on sm_100: two lanes per issue slot).  The 16-bit
// kernels below are issue-bound before they are HBM-bound (ncu r02a: layer-norm forward bf16 at 72 %
// issue-slot utilisation for 38 % of DRAM throughput), so every element-wise step works on pairs.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)),
        "l"(reinterpret_cast<const uint64_t&>(c)));
  return d;
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<const uint64_t&>(a)), "l"(reinterpret_cast<const uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float2 f2_dup(float v) { return make_float2(v, v); }

// 16-byte vector <-> NP = VE / 2 float pairs
template <typename T>
__device__ __forceinline__ void unpack_pairs(const uint4& v, float2 (&out)[VecIO<T>::N / 2]);
template <>
__device__ __forceinline__ void unpack_pairs<float>(const uint4& v, float2 (&out)[2]) {
  out[0] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
  out[1] = make_float2(__uint_as_float(v.z), __uint_as_float(v.w));
}
template <>
__device__ __forceinline__ void unpack_pairs<__half>(const uint4& v, float2 (&out)[4]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = __half22float2(h[i]);
}
template <>
__device__ __forceinline__ void unpack_pairs<__nv_bfloat16>(const uint4& v, float2 (&out)[4]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
}
template <typename T>
__device__ __forceinline__ void store_pairs(T* p, const float2 (&in)[VecIO<T>::N / 2]) {
  uint4 v;
  if constexpr (sizeof(T) == 4) {
    v = make_uint4(__float_as_uint(in[0].x), __float_as_uint(in[0].y), __float_as_uint(in[1].x),
                   __float_as_uint(in[1].y));
  } else {
    v = make_uint4(pack2<T>(in[0].x, in[0].y), pack2<T>(in[1].x, in[1].y), pack2<T>(in[2].x, in[2].y),
                   pack2<T>(in[3].x, in[3].y));
  }
  *reinterpret_cast<uint4*>(p) = v;
}
template <typename T>
__device__ __forceinline__ void load_pairs(const T* p, float2 (&out)[VecIO<T>::N / 2]) {
  unpack_pairs<T>(*reinterpret_cast<const uint4*>(p), out);
}

// ---------------------------------------------------------------------------------------
// forward kernels (vector path).  OP: 0 = softmax, 1 = rms norm, 2 = layer norm.
// Persistent grid (as many CTAs as are resident); CTA g handles row groups g, g+G, ...  Rows are
// streamed through a 3-stage cp.async ring in shared memory -- every thread copies exactly the
// vectors it later reads itself, so the ring needs no block barrier -- which keeps two row groups in
// flight per CTA while it reduces the current one; w / b stay in registers across rows when the row is
// short enough (MAXV <= 2), otherwise they are re-read from L1 / L2 per row.
// ---------------------------------------------------------------------------------------
// FULL: the row is exactly TPR * MAXV vectors (emb 4096 bf16, 4096 / 8192 Float32, ...): every "is this vector
// inside the row" predicate folds away -- these kernels are issue-bound for 16-bit rows.
template <typename T, int TPR, int MAXV, int OP, bool PERSIST, bool FULL = false>
__global__ void __launch_bounds__(kThreads)
rowwise_fwd_vec(T* __restrict__ y, float* __restrict__ stat0, float* __restrict__ stat1,
                const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b,
                int64_t emb, int64_t n, float eps, float offset) {
  constexpr int VE = VecIO<T>::N;
  constexpr int NP = VE / 2;
  constexpr int RPB = kThreads / TPR;
  constexpr bool kCacheW = PERSIST && OP != 0 && MAXV <= 2;
  __shared__ RedBuf red;
  int parity = 0;
  const int t = threadIdx.x % TPR;
  const int sub = threadIdx.x / TPR;
  const int nvec = static_cast<int>(emb / VE);
  auto in_row = [&](int vi) { return FULL || vi < nvec; };
  const float inv_n = 1.f / static_cast<float>(emb);
  const int64_t n_groups = (n + RPB - 1) / RPB;

  float2 wv[kCacheW ? MAXV : 1][NP], bv[(kCacheW && OP == 2) ? MAXV : 1][NP];
  auto load_wb = [&](int i, float2 (&wo)[NP], float2 (&bo)[NP]) {
    const int vi = t + i * TPR;
    load_pairs<T>(w + static_cast<int64_t>(vi) * VE, wo);
    if constexpr (OP == 1) {
#pragma unroll
      for (int j = 0; j < NP; ++j) wo[j] = f2_add(wo[j], f2_dup(offset));
    }
    if constexpr (OP == 2) load_pairs<T>(b + static_cast<int64_t>(vi) * VE, bo);
  };
  if constexpr (kCacheW) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (in_row(t + i * TPR)) {
        load_wb(i, wv[i], bv[OP == 2 ? i : 0]);
      } else {
#pragma unroll
        for (int j = 0; j < NP; ++j) wv[i][j] = bv[OP == 2 ? i : 0][j] = f2_dup(0.f);
      }
    }
  }

  extern __shared__ __align__(16) uint8_t ring[];  // PERSIST: [3 stages][MAXV][256 threads] x 16 B
  auto slot = [&](int st, int i) {
    return ring + ((static_cast<size_t>(st * MAXV + i) * kThreads + threadIdx.x) << 4);
  };
  auto fetch = [&](int64_t g, int st) {
    const int64_t row = g * RPB + sub;
    if (g < n_groups && row < n) {
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (in_row(vi)) cp_async16(slot(st, i), x + row * emb + static_cast<int64_t>(vi) * VE);
      }
    }
    cp_async_commit();  // always commit so the group count is uniform
  };
  if constexpr (PERSIST) {
    fetch(blockIdx.x, 0);
    fetch(static_cast<int64_t>(blockIdx.x) + gridDim.x, 1);
  }
  int stage = 0;
  // !PERSIST: one row group per CTA (grid = n_groups), the row loaded straight into registers
  for (int64_t g = blockIdx.x; g < n_groups; g += PERSIST ? gridDim.x : n_groups) {
    const int64_t row = g * RPB + sub;
    const bool live = row < n;  // TPR==256: uniform per CTA; TPR==32: uniform per warp
    if constexpr (PERSIST) {
      fetch(g + 2 * static_cast<int64_t>(gridDim.x), stage == 0 ? 2 : stage - 1);
      cp_async_wait<2>();  // everything but the two newest groups has landed
    }
    float2 xv[MAXV][NP];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (live && in_row(t + i * TPR)) {
        if constexpr (PERSIST)
          unpack_pairs<T>(*reinterpret_cast<const uint4*>(slot(stage, i)), xv[i]);
        else
          load_pairs<T>(x + row * emb + static_cast<int64_t>(t + i * TPR) * VE, xv[i]);
      } else {
#pragma unroll
        for (int j = 0; j < NP; ++j) xv[i][j] = f2_dup(OP == 0 ? -INFINITY : 0.f);
      }
    }
    stage = stage == 2 ? 0 : stage + 1;
    if (!live) continue;  // (no block barrier is used when TPR == 32)
    T* yr = y + row * emb;

    if constexpr (OP == 0) {
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) m = fmaxf(m, fmaxf(xv[i][j].x, xv[i][j].y));
      m = row_max<TPR>(m, red, parity);
      const float ml2 = (m == -INFINITY) ? 0.f : m * 1.4426950408889634f;
      const float2 l2e = f2_dup(1.4426950408889634f), nm = f2_dup(-ml2);
      float2 s2 = f2_dup(0.f);
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const float2 a = f2_fma(xv[i][j], l2e, nm);
          xv[i][j] = make_float2(fast_exp2(a.x), fast_exp2(a.y));
          s2 = f2_add(s2, xv[i][j]);
        }
      const float s = row_sum<TPR>(s2.x + s2.y, red, parity);
      const float2 inv = f2_dup(1.f / s);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (in_row(vi)) {
#pragma unroll
          for (int j = 0; j < NP; ++j) xv[i][j] = f2_mul(xv[i][j], inv);
          store_pairs<T>(yr + static_cast<int64_t>(vi) * VE, xv[i]);
        }
      }
    } else if constexpr (OP == 1) {
      float2 ss2 = f2_dup(0.f);
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) ss2 = f2_fma(xv[i][j], xv[i][j], ss2);
      const float ss = row_sum<TPR>(ss2.x + ss2.y, red, parity);
      const float rstd = rsqrtf(ss * inv_n + eps);
      if (t == 0) stat0[row] = rstd;
      const float2 r2 = f2_dup(rstd);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (in_row(vi)) {
          float2 wl[NP], bl[NP];
          if constexpr (!kCacheW) load_wb(i, wl, bl);
#pragma unroll
          for (int j = 0; j < NP; ++j) xv[i][j] = f2_mul(f2_mul(xv[i][j], r2), kCacheW ? wv[i][j] : wl[j]);
          store_pairs<T>(yr + static_cast<int64_t>(vi) * VE, xv[i]);
        }
      }
    } else {
      float2 s2 = f2_dup(0.f);
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) s2 = f2_add(s2, xv[i][j]);
      const float mu = row_sum<TPR>(s2.x + s2.y, red, parity) * inv_n;
      const float2 nmu = f2_dup(-mu);
      float2 ss2 = f2_dup(0.f);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        if (in_row(t + i * TPR)) {  // padding vectors hold 0, not mu: keep them out of the variance
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            xv[i][j] = f2_add(xv[i][j], nmu);
            ss2 = f2_fma(xv[i][j], xv[i][j], ss2);
          }
        }
      }
      const float ss = row_sum<TPR>(ss2.x + ss2.y, red, parity);
      const float rstd = rsqrtf(ss * inv_n + eps);
      if (t == 0) {
        stat0[row] = mu;
        stat1[row] = rstd;
      }
      const float2 r2 = f2_dup(rstd);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (in_row(vi)) {
          float2 wl[NP], bl[NP];
          if constexpr (!kCacheW) load_wb(i, wl, bl);
#pragma unroll
          for (int j = 0; j < NP; ++j)
            xv[i][j] = f2_fma(f2_mul(xv[i][j], r2), kCacheW ? wv[i][j] : wl[j], kCacheW ? bv[i][j] : bl[j]);
          store_pairs<T>(yr + static_cast<int64_t>(vi) * VE, xv[i]);
        }
      }
    }
  }
}

// scalar fallback: one CTA per row, any emb / alignment; re-reads the row (L1/L2 hits)
template <typename T, int OP>
__global__ void __launch_bounds__(kThreads)
rowwise_fwd_generic(T* __restrict__ y, float* __restrict__ stat0, float* __restrict__ stat1,
                    const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b,
                    int64_t emb, int64_t n, float eps, float offset) {
  __shared__ RedBuf red;
  int parity = 0;
  const int64_t row = blockIdx.x;
  const T* xr = x + row * emb;
  T* yr = y + row * emb;
  const float inv_n = 1.f / static_cast<float>(emb);
  if constexpr (OP == 0) {
    float m = -INFINITY;
    for (int64_t e = threadIdx.x; e < emb; e += kThreads) m = fmaxf(m, to_f32<T>(xr[e]));
    m = row_max<256>(m, red, parity);
    if (m == -INFINITY) m = 0.f;
    float s = 0.f;
    for (int64_t e = threadIdx.x; e < emb; e += kThreads) s += __expf(to_f32<T>(xr[e]) - m);
    s = row_sum<256>(s, red, parity);
    const float inv = 1.f / s;
    for (int64_t e = threadIdx.x; e < emb; e += kThreads)
      yr[e] = from_f32<T>(__expf(to_f32<T>(xr[e]) - m) * inv);
  } else if constexpr (OP == 1) {
    float ss = 0.f;
    for (int64_t e = threadIdx.x; e < emb; e += kThreads) {
      const float v = to_f32<T>(xr[e]);
      ss = fmaf(v, v, ss);
    }
    ss = row_sum<256>(ss, red, parity);
    const float rstd = rsqrtf(ss * inv_n + eps);
    if (threadIdx.x == 0) stat0[row] = rstd;
    for (int64_t e = threadIdx.x; e < emb; e += kThreads)
      yr[e] = from_f32<T>((to_f32<T>(w[e]) + offset) * to_f32<T>(xr[e]) * rstd);
  } else {
    float s = 0.f;
    for (int64_t e = threadIdx.x; e < emb; e += kThreads) s += to_f32<T>(xr[e]);
    s = row_sum<256>(s, red, parity);
    const float mu = s * inv_n;
    float ss = 0.f;
    for (int64_t e = threadIdx.x; e < emb; e += kThreads) {
      const float d = to_f32<T>(xr[e]) - mu;
      ss = fmaf(d, d, ss);
    }
    ss = row_sum<256>(ss, red, parity);
    const float rstd = rsqrtf(ss * inv_n + eps);
    if (threadIdx.x == 0) {
      stat0[row] = mu;
      stat1[row] = rstd;
    }
    for (int64_t e = threadIdx.x; e < emb; e += kThreads)
      yr[e] = from_f32<T>(fmaf((to_f32<T>(xr[e]) - mu) * rstd, to_f32<T>(w[e]), to_f32<T>(b[e])));
  }
}

// ---------------------------------------------------------------------------------------
// backward kernels (vector path).  Persistent grid; CTA g handles row groups g, g+G, ...
// partial layout: part0[g][emb] (dw), part1[g][emb] (db, layer norm only), fp32, one row per CTA g.
// ---------------------------------------------------------------------------------------
// NT = threads per CTA (256 in production; the 512-thread form -- half the columns, accumulators and registers
// per thread, twice the warps per SM -- is kept for the A/B recorded in launch_bwd: it lost).
template <typename T, int TPR, int MAXV, int OP, int NT = kThreads, bool FULL = false>
__global__ void __launch_bounds__(NT, (NT == 512 ? 2 : (MAXV <= 2 ? NNOP_ROWWISE_BWD_MINB : 1)))
rowwise_bwd_vec(T* __restrict__ dx, float* __restrict__ part0, float* __restrict__ part1,
                const T* __restrict__ dy, const T* __restrict__ x_or_y,
                const float* __restrict__ stat0, const float* __restrict__ stat1,
                const T* __restrict__ w, int64_t emb, int64_t n, float offset) {
  constexpr int VE = VecIO<T>::N;
  constexpr int NP = VE / 2;
  constexpr int RPB = NT / TPR;
  constexpr int NW = NT / 32;
  __shared__ RedBuf red;
  int parity = 0;
  // the partial-reduction kernel that follows may be scheduled as soon as SMs free up (it waits for this
  // grid's completion itself: griddepcontrol.wait); a no-op when nothing depends programmatically
  if constexpr (OP != 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int t = threadIdx.x % TPR;
  const int sub = threadIdx.x / TPR;
  const int nvec = static_cast<int>(emb / VE);
  auto in_row = [&](int vi) { return FULL || vi < nvec; };
  const float inv_n = 1.f / static_cast<float>(emb);
  const int64_t n_groups = (n + RPB - 1) / RPB;

  float2 wv[MAXV][NP];
  float2 acc0[MAXV][NP];
  float2 acc1[(OP == 2) ? MAXV : 1][NP];
  if constexpr (OP != 0) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = t + i * TPR;
      if (in_row(vi)) {
        load_pairs<T>(w + static_cast<int64_t>(vi) * VE, wv[i]);
      } else {
#pragma unroll
        for (int j = 0; j < NP; ++j) wv[i][j] = f2_dup(0.f);
      }
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        if (OP == 1) wv[i][j] = f2_add(wv[i][j], f2_dup(offset));
        acc0[i][j] = f2_dup(0.f);
        if (OP == 2) acc1[i][j] = f2_dup(0.f);
      }
    }
  }

  // Row groups are streamed through a 3-stage cp.async ring in shared memory.  Every thread
  // copies exactly the vectors it later reads itself, so the ring needs no block barrier, and a
  // CTA keeps two row groups of x and dy in flight while it reduces the current one.
  extern __shared__ __align__(16) uint8_t ring[];  // [3 stages][x|dy][MAXV][NT threads] x 16 B
  auto slot = [&](int st, int which, int i) {
    return ring + ((static_cast<size_t>((st * 2 + which) * MAXV + i) * NT + threadIdx.x) << 4);
  };
  auto fetch = [&](int64_t g, int st) {
    const int64_t row = g * RPB + sub;
    if (g < n_groups && row < n) {
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (in_row(vi)) {
          cp_async16(slot(st, 0, i), x_or_y + row * emb + static_cast<int64_t>(vi) * VE);
          cp_async16(slot(st, 1, i), dy + row * emb + static_cast<int64_t>(vi) * VE);
        }
      }
    }
    cp_async_commit();  // always commit so the group count is uniform
  };
  fetch(blockIdx.x, 0);
  fetch(static_cast<int64_t>(blockIdx.x) + gridDim.x, 1);
  // the row's saved statistics (rstd; mean, rstd) are fetched one row group ahead as well: read at the
  // point of use they were a dependent L2 / HBM round trip at the head of every row's chain
  auto stat = [&](const float* p, int64_t g) {
    const int64_t row = g * RPB + sub;
    return (OP != 0 && p != nullptr && g < n_groups && row < n) ? p[row] : 0.f;
  };
  float st0_next = stat(stat0, blockIdx.x), st1_next = OP == 2 ? stat(stat1, blockIdx.x) : 0.f;
  int stage = 0;
  for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int64_t row = g * RPB + sub;
    const bool live = row < n;  // TPR==256: uniform per CTA; TPR==32: uniform per warp
    fetch(g + 2 * static_cast<int64_t>(gridDim.x), stage == 0 ? 2 : stage - 1);
    const float st0_cur = st0_next, st1_cur = st1_next;
    st0_next = stat(stat0, g + gridDim.x);
    if (OP == 2) st1_next = stat(stat1, g + gridDim.x);
    cp_async_wait<2>();  // everything but the two newest groups has landed
    float2 av[MAXV][NP], dv[MAXV][NP];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = t + i * TPR;
      if (live && in_row(vi)) {
        unpack_pairs<T>(*reinterpret_cast<const uint4*>(slot(stage, 0, i)), av[i]);
        unpack_pairs<T>(*reinterpret_cast<const uint4*>(slot(stage, 1, i)), dv[i]);
      } else {
#pragma unroll
        for (int j = 0; j < NP; ++j) av[i][j] = dv[i][j] = f2_dup(0.f);
      }
    }
    stage = stage == 2 ? 0 : stage + 1;
    if (!live) continue;        // (no block barrier is used when TPR == 32)
    T* dxr = dx + row * emb;
    if constexpr (OP == 0) {
      // dx = y*dy - y*sum(y*dy)              (src/softmax.jl:70-80)
      float2 s2 = f2_dup(0.f);
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) s2 = f2_fma(av[i][j], dv[i][j], s2);
      const float2 ns = f2_dup(-row_sum<TPR, NW>(s2.x + s2.y, red, parity));
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (in_row(vi)) {
#pragma unroll
          for (int j = 0; j < NP; ++j) av[i][j] = f2_mul(av[i][j], f2_add(dv[i][j], ns));
          store_pairs<T>(dxr + static_cast<int64_t>(vi) * VE, av[i]);
        }
      }
    } else if constexpr (OP == 1) {
      // dx = r*d*(w+off) - r^3*x*sum(d*(w+off)*x)/N ; dw += d*x*r   (src/rms_norm.jl:72-101)
      const float r = st0_cur;
      float2 dd2 = f2_dup(0.f);
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) dd2 = f2_fma(f2_mul(dv[i][j], wv[i][j]), av[i][j], dd2);
      const float dd = row_sum<TPR, NW>(dd2.x + dd2.y, red, parity);
      const float2 r2 = f2_dup(r), nc = f2_dup(-(r * r * r * dd * inv_n));
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (in_row(vi)) {
          float2 o[NP];
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const float2 dr = f2_mul(dv[i][j], r2);
            acc0[i][j] = f2_fma(dr, av[i][j], acc0[i][j]);
            o[j] = f2_fma(dr, wv[i][j], f2_mul(nc, av[i][j]));
          }
          store_pairs<T>(dxr + static_cast<int64_t>(vi) * VE, o);
        }
      }
    } else {
      // xh=(x-mu)r; c1=mean(w d xh); c2=mean(w d); dx=(w d-(xh c1+c2)) r   (src/layer_norm.jl:95-136)
      // The kernel is issue-bound for 16-bit rows, so the operation count per element pair matters: 8 packed
      // ops here (xh 1, w*dy 1, two row sums 2, two accumulators 2, dx 2), w*dy kept in registers between phases.
      const float mu = st0_cur;
      const float r = st1_cur;
      const float2 r2 = f2_dup(r), nmur = f2_dup(-mu * r);
      float2 s1 = f2_dup(0.f), s2 = f2_dup(0.f);
      float2 wd[MAXV][NP];
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          av[i][j] = f2_fma(av[i][j], r2, nmur);  // xh = (x - mu) r  (padding vectors: w = d = 0 below)
          wd[i][j] = f2_mul(dv[i][j], wv[i][j]);
          s1 = f2_fma(wd[i][j], av[i][j], s1);
          s2 = f2_add(s2, wd[i][j]);
        }
      float c1 = s1.x + s1.y, c2 = s2.x + s2.y;
      row_sum2<TPR, NW>(c1, c2, red, parity);
      const float2 nc1r = f2_dup(-c1 * inv_n * r), nc2r = f2_dup(-c2 * inv_n * r);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (in_row(vi)) {
          float2 o[NP];
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            acc0[i][j] = f2_fma(dv[i][j], av[i][j], acc0[i][j]);
            acc1[i][j] = f2_add(acc1[i][j], dv[i][j]);
            o[j] = f2_fma(wd[i][j], r2, f2_fma(av[i][j], nc1r, nc2r));   // (w d - (xh c1 + c2)) r
          }
          store_pairs<T>(dxr + static_cast<int64_t>(vi) * VE, o);
        }
      }
    }
  }

  if constexpr (OP != 0 && TPR == 32) {
    // warp-per-row form: the CTA's RPB warps first add their accumulators in shared memory (the prefetch ring,
    // idle by now), so the CTA writes ONE partial row instead of RPB (r02a: 1 024 rows of emb 1 024 wrote 1 024
    // partial rows -- twice the bytes of the input -- for the second kernel to read back)
    constexpr int EP = 32 * MAXV * VE;   // padded row length in floats
    cp_async_wait<0>();
    __syncthreads();
    float* rows = reinterpret_cast<float*>(ring);   // [dw | db][sub][EP]
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = t + i * TPR;
#pragma unroll
      for (int j = 0; j < NP; j += 2) {
        *reinterpret_cast<float4*>(rows + sub * EP + vi * VE + 2 * j) =
            make_float4(acc0[i][j].x, acc0[i][j].y, acc0[i][j + 1].x, acc0[i][j + 1].y);
        if (OP == 2)
          *reinterpret_cast<float4*>(rows + (RPB + sub) * EP + vi * VE + 2 * j) =
              make_float4(acc1[i][j].x, acc1[i][j].y, acc1[i][j + 1].x, acc1[i][j + 1].y);
      }
    }
    __syncthreads();
    for (int c4 = threadIdx.x; c4 < (OP == 2 ? 2 : 1) * (EP / 4); c4 += NT) {
      const int which = c4 / (EP / 4), col = (c4 - which * (EP / 4)) * 4;
      if (col >= emb) continue;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < RPB; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(rows + (which * RPB + r) * EP + col);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      *reinterpret_cast<float4*>((which ? part1 : part0) + static_cast<int64_t>(blockIdx.x) * emb + col) = a;
    }
  } else if constexpr (OP != 0) {
    // one partial row per CTA: index blockIdx.x
    const int64_t prow = static_cast<int64_t>(blockIdx.x) * RPB + sub;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = t + i * TPR;
      if (in_row(vi)) {
#pragma unroll
        for (int j = 0; j < NP; j += 2) {
          *reinterpret_cast<float4*>(part0 + prow * emb + static_cast<int64_t>(vi) * VE + 2 * j) =
              make_float4(acc0[i][j].x, acc0[i][j].y, acc0[i][j + 1].x, acc0[i][j + 1].y);
          if (OP == 2)
            *reinterpret_cast<float4*>(part1 + prow * emb + static_cast<int64_t>(vi) * VE + 2 * j) =
                make_float4(acc1[i][j].x, acc1[i][j].y, acc1[i][j + 1].x, acc1[i][j + 1].y);
        }
      }
    }
  }
}

// scalar fallback backward: CTA g handles rows g, g+G, ...; thread owns columns tid, tid+256..
template <typename T, int OP>
__global__ void __launch_bounds__(kThreads)
rowwise_bwd_generic(T* __restrict__ dx, float* __restrict__ part0, float* __restrict__ part1,
                    const T* __restrict__ dy, const T* __restrict__ x_or_y,
                    const float* __restrict__ stat0, const float* __restrict__ stat1,
                    const T* __restrict__ w, int64_t emb, int64_t n, float offset) {
  __shared__ RedBuf red;
  int parity = 0;
  const float inv_n = 1.f / static_cast<float>(emb);
  float* p0 = part0 ? part0 + static_cast<int64_t>(blockIdx.x) * emb : nullptr;
  float* p1 = part1 ? part1 + static_cast<int64_t>(blockIdx.x) * emb : nullptr;
  if (OP != 0) {
    for (int64_t e = threadIdx.x; e < emb; e += kThreads) {
      p0[e] = 0.f;
      if (OP == 2) p1[e] = 0.f;
    }
  }
  for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
    const T* ar = x_or_y + row * emb;
    const T* dr = dy + row * emb;
    T* dxr = dx + row * emb;
    if constexpr (OP == 0) {
      float s = 0.f;
      for (int64_t e = threadIdx.x; e < emb; e += kThreads)
        s = fmaf(to_f32<T>(ar[e]), to_f32<T>(dr[e]), s);
      s = row_sum<256>(s, red, parity);
      for (int64_t e = threadIdx.x; e < emb; e += kThreads)
        dxr[e] = from_f32<T>(to_f32<T>(ar[e]) * (to_f32<T>(dr[e]) - s));
    } else if constexpr (OP == 1) {
      const float r = stat0[row];
      float dd = 0.f;
      for (int64_t e = threadIdx.x; e < emb; e += kThreads)
        dd = fmaf(to_f32<T>(dr[e]) * (to_f32<T>(w[e]) + offset), to_f32<T>(ar[e]), dd);
      dd = row_sum<256>(dd, red, parity);
      const float c = r * r * r * dd * inv_n;
      for (int64_t e = threadIdx.x; e < emb; e += kThreads) {
        const float d = to_f32<T>(dr[e]), xv = to_f32<T>(ar[e]);
        p0[e] = fmaf(d * xv, r, p0[e]);
        dxr[e] = from_f32<T>(fmaf(r * d, to_f32<T>(w[e]) + offset, -c * xv));
      }
    } else {
      const float mu = stat0[row], r = stat1[row];
      float s1 = 0.f, s2 = 0.f;
      for (int64_t e = threadIdx.x; e < emb; e += kThreads) {
        const float xh = (to_f32<T>(ar[e]) - mu) * r;
        const float wd = to_f32<T>(dr[e]) * to_f32<T>(w[e]);
        s1 = fmaf(wd, xh, s1);
        s2 += wd;
      }
      row_sum2<256>(s1, s2, red, parity);
      s1 *= inv_n;
      s2 *= inv_n;
      for (int64_t e = threadIdx.x; e < emb; e += kThreads) {
        const float xh = (to_f32<T>(ar[e]) - mu) * r;
        const float d = to_f32<T>(dr[e]);
        p0[e] = fmaf(d, xh, p0[e]);
        p1[e] += d;
        dxr[e] = from_f32<T>((d * to_f32<T>(w[e]) - fmaf(xh, s1, s2)) * r);
      }
    }
  }
}

// out{0,1}[e] = sum_g part{0,1}[g][e]; blockDim (8, 32) = 8 float4 column groups x 32 row walkers, grid
// (ceil(emb/32), 1 or 2): blockIdx.y picks dw / db.  The kernel runs behind a 40 us operation on data that sits
// in L2, so what it costs is (i) dependent L2 round trips and (ii) waves: r01 walked one chain of n_part / 8 loads
// (8.5-11 us); r02a had 1 024-thread CTAs, two per SM -- layer norm at emb 8 192 needed 512 of them, two waves, and
// cost 9 us of a 49 us backward.  Now 256-thread CTAs (all resident at once), every walker issues all of its
// (up to kBatch) 16-byte loads before the first add, and a warp reads four full 128-byte row segments.
template <typename TO>
__global__ void __launch_bounds__(256)
reduce_partials(TO* __restrict__ out0, TO* __restrict__ out1, const float* __restrict__ part0,
                const float* __restrict__ part1, int64_t n_part, int64_t emb) {
  constexpr int kBatch = 10;
  __shared__ float4 sm[32][9];
  const float* part = blockIdx.y ? part1 : part0;
  TO* out = blockIdx.y ? out1 : out0;
  const int64_t e = static_cast<int64_t>(blockIdx.x) * 32 + threadIdx.x * 4;   // emb % 4 == 0 (vector paths) or scalar tail below
  // launched with programmatic stream serialisation: the CTAs are resident while the backward kernel drains
  // and only wait here for its memory to be visible (takes the launch latency off a 40 us operation)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vec = (emb & 3) == 0;
  if (vec && e < emb) {
    for (int64_t g0 = threadIdx.y; g0 < n_part; g0 += 32 * kBatch) {
      float4 v[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int64_t g = g0 + 32 * k;
        v[k] = g < n_part ? __ldcg(reinterpret_cast<const float4*>(part + g * emb + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < kBatch; ++k) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
    }
  } else if (!vec) {
    // (generic kernels with emb not a multiple of 4: scalar columns, same walk)
    float* sp = &s.x;
    for (int c = 0; c < 4; ++c)
      if (e + c < emb)
        for (int64_t g = threadIdx.y; g < n_part; g += 32) sp[c] += part[g * emb + e + c];
  }
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float4 v = sm[i][threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (e + c < emb) out[e + c] = from_f32<TO>(tv[c]);
  }
}

// ---------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// number of partial rows the backward kernels write for (emb, n); must match the launch
struct BwdPlan {
  int mode;  // 0 generic, 1 warp-per-row, 2 cta-per-row
  int grid;
  int64_t n_part;
};
template <typename T>
BwdPlan bwd_plan(int64_t emb, int64_t n, bool all_aligned, bool softmax) {
  constexpr int VE = 16 / sizeof(T);
  BwdPlan p;
  const int64_t nvec = emb / VE;
  const int max_ctas = sm_count() * 4;
  constexpr int MAXV = 4;
  if (all_aligned && emb % VE == 0 && nvec <= 32 * MAXV) {
    p.mode = 1;
    const int64_t groups = (n + 7) / 8;
    p.grid = static_cast<int>(groups < max_ctas ? groups : max_ctas);
    p.n_part = static_cast<int64_t>(p.grid) * 8;
  } else if (all_aligned && emb % VE == 0 && nvec <= 256 * (softmax ? 8 : MAXV)) {
    p.mode = 2;
    p.grid = static_cast<int>(n < max_ctas ? n : max_ctas);
    p.n_part = p.grid;
  } else {
    p.mode = 0;
    p.grid = static_cast<int>(n < max_ctas ? n : max_ctas);
    p.n_part = p.grid;
  }
  if (p.grid < 1) p.grid = 1;
  return p;
}

template <typename T, int OP>
int launch_fwd(void* y, float* s0, float* s1, const void* x, const void* w, const void* b,
               int64_t emb, int64_t n, float eps, float offset, cudaStream_t st) {
  constexpr int VE = 16 / sizeof(T);
  if (n == 0 || emb == 0) return NNOP_OK;
  const int64_t nvec = emb / VE;
  const bool al = emb % VE == 0 && aligned16(y) && aligned16(x) && (OP == 0 || aligned16(w)) &&
                  (OP != 2 || aligned16(b));
  T* yy = static_cast<T*>(y);
  const T* xx = static_cast<const T*>(x);
  const T* ww = static_cast<const T*>(w);
  const T* bb = static_cast<const T*>(b);
  // smallest register cache (MAXV vectors per thread) that holds the row.  Two launch shapes, chosen
  // by how many waves the rows make (measured, profiles/r02_perf_rowwise.txt): one row group per CTA
  // when all of them are resident at once (no CTA then does two rows while another does one) or when
  // there are so many waves that the tail does not matter and full occupancy (more warps per SM) wins;
  // in between, a persistent grid of resident CTAs with a cp.async prefetch ring and w / b kept in
  // registers (no wave quantisation, no per-row reload of the weights).
  auto run = [&](auto kern_p, auto kern_1, int rows_per_cta, int maxv) {
    const int smem = 3 * maxv * kThreads * 16;
    if (smem > 32 * 1024) cudaFuncSetAttribute(kern_p, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ_p = 1, occ_1 = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_p, kern_p, kThreads, smem) != cudaSuccess || occ_p < 1)
      occ_p = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_1, kern_1, kThreads, 0) != cudaSuccess || occ_1 < 1)
      occ_1 = 1;
    if (occ_p > 8) occ_p = 8;
    const int64_t groups = (n + rows_per_cta - 1) / rows_per_cta;
    const int64_t cap_p = static_cast<int64_t>(sm_count()) * occ_p, cap_1 = static_cast<int64_t>(sm_count()) * occ_1;
    // (between one wave and three persistent rounds the persistent grid's round quantisation is the worse of
    // the two: 1 024 softmax rows on 592 CTAs are 2 rows for most and 1 for the rest)
    if (groups <= cap_1 || groups < 3 * cap_p || groups >= 16 * cap_p) {
      kern_1<<<static_cast<unsigned>(groups), kThreads, 0, st>>>(yy, s0, s1, xx, ww, bb, emb, n, eps, offset);
    } else {
      kern_p<<<static_cast<unsigned>(cap_p), kThreads, smem, st>>>(yy, s0, s1, xx, ww, bb, emb, n, eps, offset);
    }
  };
#define NNOP_FWD(TPR_, MAXV_, RPC_)                                                                          \
  do {                                                                                                       \
    if (nvec == TPR_ * MAXV_)                                                                                \
      run(rowwise_fwd_vec<T, TPR_, MAXV_, OP, true, true>, rowwise_fwd_vec<T, TPR_, MAXV_, OP, false, true>, \
          RPC_, MAXV_);                                                                                      \
    else                                                                                                     \
      run(rowwise_fwd_vec<T, TPR_, MAXV_, OP, true>, rowwise_fwd_vec<T, TPR_, MAXV_, OP, false>, RPC_, MAXV_); \
  } while (0)
  if (al && nvec <= 32 * 8) {
    if (nvec <= 32) NNOP_FWD(32, 1, 8);
    else if (nvec <= 64) NNOP_FWD(32, 2, 8);
    else if (nvec <= 128) NNOP_FWD(32, 4, 8);
    else NNOP_FWD(32, 8, 8);
  } else if (al && nvec <= 256 * 8) {
    if (nvec <= 512) NNOP_FWD(256, 2, 1);
    else if (nvec <= 1024) NNOP_FWD(256, 4, 1);
    else NNOP_FWD(256, 8, 1);
#undef NNOP_FWD
  } else {
    rowwise_fwd_generic<T, OP><<<static_cast<unsigned>(n), kThreads, 0, st>>>(
        yy, s0, s1, xx, ww, bb, emb, n, eps, offset);
  }
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

template <typename T, int OP, typename TW>
int launch_bwd(void* dx, TW* dw, TW* db, const void* dy, const void* a, const float* s0,
               const float* s1, const void* w, int64_t emb, int64_t n, float offset, void* ws,
               size_t ws_bytes, cudaStream_t st) {
  constexpr int VE = 16 / sizeof(T);
  if (n == 0 || emb == 0) return NNOP_OK;
  const bool al = aligned16(dx) && aligned16(dy) && aligned16(a) && (OP == 0 || aligned16(w)) &&
                  (OP == 0 || aligned16(ws));
  const BwdPlan plan = bwd_plan<T>(emb, n, al, OP == 0);
  float* p0 = nullptr;
  float* p1 = nullptr;
  if (OP != 0) {
    const size_t need = static_cast<size_t>(plan.n_part) * emb * sizeof(float) * (OP == 2 ? 2 : 1);
    if (ws == nullptr || ws_bytes < need)
      return fail(NNOP_ERR_WORKSPACE, "norm backward needs a %zu-byte workspace, got %zu", need,
                  ws_bytes);
    p0 = static_cast<float*>(ws);
    p1 = (OP == 2) ? p0 + plan.n_part * emb : nullptr;
  }
  T* dxx = static_cast<T*>(dx);
  const T* dyy = static_cast<const T*>(dy);
  const T* aa = static_cast<const T*>(a);
  const T* ww = static_cast<const T*>(w);
  const int64_t nvec = emb / VE;
  // persistent grid = what is actually resident (occupancy x SMs, at most 4 per SM), so every
  // CTA streams rows back to back with its prefetch pipeline full
  int grid = plan.grid;
  int64_t n_part = plan.n_part;
  auto run = [&](auto kern, int rows_per_cta, int maxv, int nt = kThreads) {
    const int smem = maxv > 0 ? 3 * 2 * maxv * nt * 16 : 0;
    if (smem > 32 * 1024)  // static smem (reduction scratch) counts against the 48 KB default too
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem) != cudaSuccess || occ < 1)
      occ = 1;
    if (occ > 4) occ = 4;
    const int64_t groups = (n + rows_per_cta - 1) / rows_per_cta;
    const int64_t cap = static_cast<int64_t>(sm_count()) * occ;
    grid = static_cast<int>(groups < cap ? groups : cap);
    n_part = grid;   // one partial row per CTA (the warp-per-row kernels add their sub-rows in shared memory)
    kern<<<grid, nt, smem, st>>>(dxx, p0, p1, dyy, aa, s0, s1, ww, emb, n, offset);
  };
  if (plan.mode == 1) {
    if (nvec <= 32) run(rowwise_bwd_vec<T, 32, 1, OP>, 8, 1);
    else if (nvec <= 64) run(rowwise_bwd_vec<T, 32, 2, OP>, 8, 2);
    else run(rowwise_bwd_vec<T, 32, 4, OP>, 8, 4);
  } else if (plan.mode == 2) {
    // (512-thread CTAs -- rowwise_bwd_vec<T, 512, MAXV / 2, OP, 512>: half the registers, twice the warps per
    // SM -- were measured SLOWER for rows of 257..1024 vectors: layer-norm backward bf16 at emb 4096 59.6 vs
    // 52.4 us at 8192 rows, 404 vs 350 us at 65536.  The kernel is bound by instruction issue (10 packed FP
    // ops per element pair against 6 for RMS norm, which runs at 95 %), not by latency, so more warps only add
    // barrier width.  profiles/r02_perf_rowwise_nt512.txt)
    // full-row specialisation (no "vector inside the row" predicates): layer norm only -- measured +4 % at C3 and
    // 89 -> 95 % of copy bandwidth at 65 536 rows there, but -18 % for RMS norm (profiles/r02_perf_rowwise_full.txt)
    constexpr bool kFullBwd = OP == 2;
    if (kFullBwd && nvec == 256) run(rowwise_bwd_vec<T, 256, 1, OP, kThreads, kFullBwd>, 1, 1);
    else if (kFullBwd && nvec == 512) run(rowwise_bwd_vec<T, 256, 2, OP, kThreads, kFullBwd>, 1, 2);
    else if (kFullBwd && nvec == 1024) run(rowwise_bwd_vec<T, 256, 4, OP, kThreads, kFullBwd>, 1, 4);
    else if (nvec <= 256) run(rowwise_bwd_vec<T, 256, 1, OP>, 1, 1);
    else if (nvec <= 512) run(rowwise_bwd_vec<T, 256, 2, OP>, 1, 2);
    else if (nvec <= 1024) run(rowwise_bwd_vec<T, 256, 4, OP>, 1, 4);
    else if constexpr (OP == 0) run(rowwise_bwd_vec<T, 256, 8, OP>, 1, 8);
  } else {
    run(rowwise_bwd_generic<T, OP>, 1, 0);
  }
  NNOP_LAUNCH_CHECK();
  if (OP != 0) {
    // every CTA writes its partial row (zeros if it never saw a live row)
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(static_cast<unsigned>((emb + 31) / 32), OP == 2 ? 2 : 1);
    cfg.blockDim = dim3(8, 32);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const float* cp0 = p0;
    const float* cp1 = p1;
    NNOP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, reduce_partials<TW>, dw, db, cp0, cp1, n_part, emb));
    NNOP_LAUNCH_CHECK();
  }
  return NNOP_OK;
}

}  // namespace
}  // namespace nnop

using namespace nnop;

#define NNOP_DISPATCH_DTYPE(dtype, ...)                                        \
  switch (dtype) {                                                             \
    case NNOP_F32: { using T = float; __VA_ARGS__; }                           \
    case NNOP_F16: { using T = __half; __VA_ARGS__; }                          \
    case NNOP_BF16: { using T = __nv_bfloat16; __VA_ARGS__; }                  \
    default: return fail(NNOP_ERR_DTYPE, "unknown dtype code %d", dtype);      \
  }

extern "C" int nnop_softmax_fwd(void* y, const void* x, int dtype, int64_t N, int64_t cols,
                                void* stream) {
  clear_error();
  if (N < 0 || cols < 0) return fail(NNOP_ERR_SHAPE, "negative size");
  if ((!y || !x) && N * cols > 0) return fail(NNOP_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NNOP_DISPATCH_DTYPE(dtype, return (launch_fwd<T, 0>(y, nullptr, nullptr, x, nullptr, nullptr, N,
                                                      cols, 0.f, 0.f, st)));
}

extern "C" int nnop_softmax_bwd(void* dx, const void* dy, const void* y, int dtype, int64_t N,
                                int64_t cols, void* stream) {
  clear_error();
  if (N < 0 || cols < 0) return fail(NNOP_ERR_SHAPE, "negative size");
  if ((!dx || !dy || !y) && N * cols > 0) return fail(NNOP_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NNOP_DISPATCH_DTYPE(dtype, return (launch_bwd<T, 0, float>(dx, nullptr, nullptr, dy, y, nullptr,
                                                             nullptr, nullptr, N, cols, 0.f,
                                                             nullptr, 0, st)));
}

extern "C" int nnop_rms_norm_fwd(void* y, float* rstd, const void* x, const void* w, int dtype,
                                 int64_t emb, int64_t n, float eps, float offset, void* stream) {
  clear_error();
  if (emb < 0 || n < 0) return fail(NNOP_ERR_SHAPE, "negative size");
  if ((!y || !rstd || !x || !w) && emb * n > 0) return fail(NNOP_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NNOP_DISPATCH_DTYPE(dtype, return (launch_fwd<T, 1>(y, rstd, nullptr, x, w, nullptr, emb, n, eps,
                                                      offset, st)));
}

extern "C" size_t nnop_norm_bwd_workspace_bytes(int64_t emb, int64_t n) {
  // upper bound over every plan: 8 partial rows per CTA, 4 CTAs per SM, dw + db
  if (emb <= 0 || n <= 0) return 0;
  int64_t rows = static_cast<int64_t>(sm_count()) * 4 * 8;
  const int64_t cap = ((n + 7) / 8) * 8;
  if (rows > cap) rows = cap;
  return static_cast<size_t>(rows) * static_cast<size_t>(emb) * sizeof(float) * 2;
}

extern "C" int nnop_rms_norm_bwd(void* dx, float* dw_f32, const void* dy, const float* rstd,
                                 const void* x, const void* w, int dtype, int64_t emb, int64_t n,
                                 float offset, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  clear_error();
  if (emb < 0 || n < 0) return fail(NNOP_ERR_SHAPE, "negative size");
  if ((!dx || !dw_f32 || !dy || !rstd || !x || !w) && emb * n > 0)
    return fail(NNOP_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NNOP_DISPATCH_DTYPE(dtype, return (launch_bwd<T, 1, float>(dx, dw_f32, nullptr, dy, x, rstd,
                                                             nullptr, w, emb, n, offset, workspace,
                                                             workspace_bytes, st)));
}

extern "C" int nnop_layer_norm_fwd(void* y, float* mean, float* rstd, const void* x, const void* w,
                                   const void* b, int dtype, int64_t emb, int64_t n, float eps,
                                   void* stream) {
  clear_error();
  if (emb < 0 || n < 0) return fail(NNOP_ERR_SHAPE, "negative size");
  if ((!y || !mean || !rstd || !x || !w || !b) && emb * n > 0)
    return fail(NNOP_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NNOP_DISPATCH_DTYPE(dtype, return (launch_fwd<T, 2>(y, mean, rstd, x, w, b, emb, n, eps, 0.f, st)));
}

extern "C" int nnop_layer_norm_bwd(void* dx, void* dw, void* db, const void* dy, const float* mean,
                                   const float* rstd, const void* x, const void* w, int dtype,
                                   int64_t emb, int64_t n, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  clear_error();
  if (emb < 0 || n < 0) return fail(NNOP_ERR_SHAPE, "negative size");
  if ((!dx || !dw || !db || !dy || !mean || !rstd || !x || !w) && emb * n > 0)
    return fail(NNOP_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NNOP_DISPATCH_DTYPE(dtype, return (launch_bwd<T, 2, T>(dx, static_cast<T*>(dw),
                                                         static_cast<T*>(db), dy, x, mean, rstd, w,
                                                         emb, n, 0.f, workspace, workspace_bytes,
                                                         st)));
}
