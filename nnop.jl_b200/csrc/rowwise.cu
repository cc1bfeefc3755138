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
using RedBuf = float[2][2][kThreads / 32];
template <int TPR>
__device__ __forceinline__ float row_sum(float v, RedBuf& red, int& parity) {
  v = warp_sum(v);
  if constexpr (TPR == 32) return v;
  float* r = red[parity][0];
  parity ^= 1;
  if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) t += r[i];
  return t;
}
// two sums in one hop (layer-norm backward: mean(w d xh) and mean(w d))
template <int TPR>
__device__ __forceinline__ void row_sum2(float& a, float& b, RedBuf& red, int& parity) {
  a = warp_sum(a);
  b = warp_sum(b);
  if constexpr (TPR == 32) return;
  float(*r)[kThreads / 32] = red[parity];
  parity ^= 1;
  if ((threadIdx.x & 31) == 0) {
    r[0][threadIdx.x >> 5] = a;
    r[1][threadIdx.x >> 5] = b;
  }
  __syncthreads();
  float ta = 0.f, tb = 0.f;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) {
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
// forward kernels (vector path).  MAXV vectors per thread are cached in registers.
// OP: 0 = softmax, 1 = rms norm, 2 = layer norm
// ---------------------------------------------------------------------------------------
template <typename T, int TPR, int MAXV, int OP>
__global__ void __launch_bounds__(kThreads)
rowwise_fwd_vec(T* __restrict__ y, float* __restrict__ stat0, float* __restrict__ stat1,
                const T* __restrict__ x, const T* __restrict__ w, const T* __restrict__ b,
                int64_t emb, int64_t n, float eps, float offset) {
  constexpr int VE = VecIO<T>::N;
  constexpr int RPB = kThreads / TPR;
  __shared__ RedBuf red;
  int parity = 0;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * RPB + threadIdx.x / TPR;
  if (row >= n) return;  // TPR==256: whole CTA exits together; TPR==32: whole warp
  const int t = threadIdx.x % TPR;
  const int nvec = static_cast<int>(emb / VE);
  const T* xr = x + row * emb;
  T* yr = y + row * emb;

  float xv[MAXV][VE];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = t + i * TPR;
    if (vi < nvec) {
      load_vec<T>(xr + static_cast<int64_t>(vi) * VE, xv[i]);
    } else {
#pragma unroll
      for (int j = 0; j < VE; ++j) xv[i][j] = (OP == 0) ? -INFINITY : 0.f;
    }
  }
  const float inv_n = 1.f / static_cast<float>(emb);

  if constexpr (OP == 0) {
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
#pragma unroll
      for (int j = 0; j < VE; ++j) m = fmaxf(m, xv[i][j]);
    m = row_max<TPR>(m, red, parity);
    const float ml2 = (m == -INFINITY) ? 0.f : m * 1.4426950408889634f;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
#pragma unroll
      for (int j = 0; j < VE; ++j) {
        xv[i][j] = fast_exp2(fmaf(xv[i][j], 1.4426950408889634f, -ml2));
        s += xv[i][j];
      }
    s = row_sum<TPR>(s, red, parity);
    const float inv = 1.f / s;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = t + i * TPR;
      if (vi < nvec) {
#pragma unroll
        for (int j = 0; j < VE; ++j) xv[i][j] *= inv;
        store_vec<T>(yr + static_cast<int64_t>(vi) * VE, xv[i]);
      }
    }
  } else if constexpr (OP == 1) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
#pragma unroll
      for (int j = 0; j < VE; ++j) ss = fmaf(xv[i][j], xv[i][j], ss);
    ss = row_sum<TPR>(ss, red, parity);
    const float rstd = rsqrtf(ss * inv_n + eps);
    if (t == 0) stat0[row] = rstd;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = t + i * TPR;
      if (vi < nvec) {
        float wv[VE];
        load_vec<T>(w + static_cast<int64_t>(vi) * VE, wv);
#pragma unroll
        for (int j = 0; j < VE; ++j) xv[i][j] = (wv[j] + offset) * xv[i][j] * rstd;
        store_vec<T>(yr + static_cast<int64_t>(vi) * VE, xv[i]);
      }
    }
  } else {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
#pragma unroll
      for (int j = 0; j < VE; ++j) s += xv[i][j];
    s = row_sum<TPR>(s, red, parity);
    const float mu = s * inv_n;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = t + i * TPR;
      if (vi < nvec) {
#pragma unroll
        for (int j = 0; j < VE; ++j) {
          const float d = xv[i][j] - mu;
          ss = fmaf(d, d, ss);
        }
      }
    }
    ss = row_sum<TPR>(ss, red, parity);
    const float rstd = rsqrtf(ss * inv_n + eps);
    if (t == 0) {
      stat0[row] = mu;
      stat1[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = t + i * TPR;
      if (vi < nvec) {
        float wv[VE], bv[VE];
        load_vec<T>(w + static_cast<int64_t>(vi) * VE, wv);
        load_vec<T>(b + static_cast<int64_t>(vi) * VE, bv);
#pragma unroll
        for (int j = 0; j < VE; ++j) xv[i][j] = fmaf((xv[i][j] - mu) * rstd, wv[j], bv[j]);
        store_vec<T>(yr + static_cast<int64_t>(vi) * VE, xv[i]);
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
// partial layout: part0[g][emb] (dw), part1[g][emb] (db, layer norm only), fp32.
// ---------------------------------------------------------------------------------------
template <typename T, int TPR, int MAXV, int OP>
__global__ void __launch_bounds__(kThreads)
rowwise_bwd_vec(T* __restrict__ dx, float* __restrict__ part0, float* __restrict__ part1,
                const T* __restrict__ dy, const T* __restrict__ x_or_y,
                const float* __restrict__ stat0, const float* __restrict__ stat1,
                const T* __restrict__ w, int64_t emb, int64_t n, float offset) {
  constexpr int VE = VecIO<T>::N;
  constexpr int RPB = kThreads / TPR;
  __shared__ RedBuf red;
  int parity = 0;
  const int t = threadIdx.x % TPR;
  const int sub = threadIdx.x / TPR;
  const int nvec = static_cast<int>(emb / VE);
  const float inv_n = 1.f / static_cast<float>(emb);
  const int64_t n_groups = (n + RPB - 1) / RPB;

  float wv[MAXV][VE];
  float acc0[MAXV][VE];
  float acc1[(OP == 2) ? MAXV : 1][VE];
  if constexpr (OP != 0) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = t + i * TPR;
      if (vi < nvec) {
        load_vec<T>(w + static_cast<int64_t>(vi) * VE, wv[i]);
      } else {
#pragma unroll
        for (int j = 0; j < VE; ++j) wv[i][j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < VE; ++j) {
        if (OP == 1) wv[i][j] += offset;
        acc0[i][j] = 0.f;
        if (OP == 2) acc1[i][j] = 0.f;
      }
    }
  }

  // Row groups are streamed through a 3-stage cp.async ring in shared memory.  Every thread
  // copies exactly the vectors it later reads itself, so the ring needs no block barrier, and a
  // CTA keeps two row groups of x and dy in flight while it reduces the current one.
  extern __shared__ __align__(16) uint8_t ring[];  // [3 stages][x|dy][MAXV][256 threads] x 16 B
  auto slot = [&](int st, int which, int i) {
    return ring + ((static_cast<size_t>((st * 2 + which) * MAXV + i) * kThreads + threadIdx.x) << 4);
  };
  auto fetch = [&](int64_t g, int st) {
    const int64_t row = g * RPB + sub;
    if (g < n_groups && row < n) {
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (vi < nvec) {
          cp_async16(slot(st, 0, i), x_or_y + row * emb + static_cast<int64_t>(vi) * VE);
          cp_async16(slot(st, 1, i), dy + row * emb + static_cast<int64_t>(vi) * VE);
        }
      }
    }
    cp_async_commit();  // always commit so the group count is uniform
  };
  fetch(blockIdx.x, 0);
  fetch(static_cast<int64_t>(blockIdx.x) + gridDim.x, 1);
  int stage = 0;
  for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const int64_t row = g * RPB + sub;
    const bool live = row < n;  // TPR==256: uniform per CTA; TPR==32: uniform per warp
    fetch(g + 2 * static_cast<int64_t>(gridDim.x), stage == 0 ? 2 : stage - 1);
    cp_async_wait<2>();  // everything but the two newest groups has landed
    float av[MAXV][VE], dv[MAXV][VE];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = t + i * TPR;
      if (live && vi < nvec) {
        unpack_raw<T>(*reinterpret_cast<const uint4*>(slot(stage, 0, i)), av[i]);
        unpack_raw<T>(*reinterpret_cast<const uint4*>(slot(stage, 1, i)), dv[i]);
      } else {
#pragma unroll
        for (int j = 0; j < VE; ++j) av[i][j] = dv[i][j] = 0.f;
      }
    }
    stage = stage == 2 ? 0 : stage + 1;
    if (!live) continue;        // (no block barrier is used when TPR == 32)
    T* dxr = dx + row * emb;
    if constexpr (OP == 0) {
      // dx = y*dy - y*sum(y*dy)              (src/softmax.jl:70-80)
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < VE; ++j) s = fmaf(av[i][j], dv[i][j], s);
      s = row_sum<TPR>(s, red, parity);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (vi < nvec) {
#pragma unroll
          for (int j = 0; j < VE; ++j) av[i][j] = av[i][j] * (dv[i][j] - s);
          store_vec<T>(dxr + static_cast<int64_t>(vi) * VE, av[i]);
        }
      }
    } else if constexpr (OP == 1) {
      // dx = r*d*(w+off) - r^3*x*sum(d*(w+off)*x)/N ; dw += d*x*r   (src/rms_norm.jl:72-101)
      const float r = stat0[row];
      float dd = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < VE; ++j) dd = fmaf(dv[i][j] * wv[i][j], av[i][j], dd);
      dd = row_sum<TPR>(dd, red, parity);
      const float c = r * r * r * dd * inv_n;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (vi < nvec) {
          float o[VE];
#pragma unroll
          for (int j = 0; j < VE; ++j) {
            acc0[i][j] = fmaf(dv[i][j] * av[i][j], r, acc0[i][j]);
            o[j] = fmaf(r * dv[i][j], wv[i][j], -c * av[i][j]);
          }
          store_vec<T>(dxr + static_cast<int64_t>(vi) * VE, o);
        }
      }
    } else {
      // xh=(x-mu)r; c1=mean(w d xh); c2=mean(w d); dx=(w d-(xh c1+c2)) r   (src/layer_norm.jl:95-136)
      const float mu = stat0[row];
      const float r = stat1[row];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < VE; ++j) {
          av[i][j] = (av[i][j] - mu) * r;  // xh
          const float wd = dv[i][j] * wv[i][j];
          s1 = fmaf(wd, av[i][j], s1);
          s2 += wd;
        }
      row_sum2<TPR>(s1, s2, red, parity);
      s1 *= inv_n;
      s2 *= inv_n;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int vi = t + i * TPR;
        if (vi < nvec) {
          float o[VE];
#pragma unroll
          for (int j = 0; j < VE; ++j) {
            acc0[i][j] = fmaf(dv[i][j], av[i][j], acc0[i][j]);
            acc1[i][j] += dv[i][j];
            o[j] = (dv[i][j] * wv[i][j] - fmaf(av[i][j], s1, s2)) * r;
          }
          store_vec<T>(dxr + static_cast<int64_t>(vi) * VE, o);
        }
      }
    }
  }

  if constexpr (OP != 0) {
    // one partial row per (CTA, sub-row): index blockIdx.x * RPB + sub
    const int64_t prow = static_cast<int64_t>(blockIdx.x) * RPB + sub;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = t + i * TPR;
      if (vi < nvec) {
#pragma unroll
        for (int j = 0; j < VE; j += 4) {
          *reinterpret_cast<float4*>(part0 + prow * emb + static_cast<int64_t>(vi) * VE + j) =
              make_float4(acc0[i][j], acc0[i][j + 1], acc0[i][j + 2], acc0[i][j + 3]);
          if (OP == 2)
            *reinterpret_cast<float4*>(part1 + prow * emb + static_cast<int64_t>(vi) * VE + j) =
                make_float4(acc1[i][j], acc1[i][j + 1], acc1[i][j + 2], acc1[i][j + 3]);
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

// out[e] = sum_g part[g][e]; blockDim (32, 8), grid ceil(emb/32)
template <typename TO>
__global__ void reduce_partials(TO* __restrict__ out, const float* __restrict__ part,
                                int64_t n_part, int64_t emb) {
  __shared__ float sm[8][33];
  const int64_t e = static_cast<int64_t>(blockIdx.x) * 32 + threadIdx.x;
  float s = 0.f;
  if (e < emb)
    for (int64_t g = threadIdx.y; g < n_part; g += 8) s += part[g * emb + e];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && e < emb) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    out[e] = from_f32<TO>(t);
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
  // smallest register cache (MAXV vectors per thread) that holds the row: fewer registers ->
  // more resident CTAs -> more bytes in flight
#define NNOP_FWD_LAUNCH(TPR_, MAXV_, GRID_)                                                    \
  rowwise_fwd_vec<T, TPR_, MAXV_, OP><<<static_cast<unsigned>(GRID_), kThreads, 0, st>>>(      \
      yy, s0, s1, xx, ww, bb, emb, n, eps, offset)
  if (al && nvec <= 32 * 8) {
    const int64_t grid = (n + 7) / 8;
    if (nvec <= 32) NNOP_FWD_LAUNCH(32, 1, grid);
    else if (nvec <= 64) NNOP_FWD_LAUNCH(32, 2, grid);
    else if (nvec <= 128) NNOP_FWD_LAUNCH(32, 4, grid);
    else NNOP_FWD_LAUNCH(32, 8, grid);
  } else if (al && nvec <= 256 * 8) {
    if (nvec <= 512) NNOP_FWD_LAUNCH(256, 2, n);
    else if (nvec <= 1024) NNOP_FWD_LAUNCH(256, 4, n);
    else NNOP_FWD_LAUNCH(256, 8, n);
  } else {
    rowwise_fwd_generic<T, OP><<<static_cast<unsigned>(n), kThreads, 0, st>>>(
        yy, s0, s1, xx, ww, bb, emb, n, eps, offset);
  }
#undef NNOP_FWD_LAUNCH
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
  auto run = [&](auto kern, int rows_per_cta, int maxv) {
    const int smem = maxv > 0 ? 3 * 2 * maxv * kThreads * 16 : 0;
    if (smem > 32 * 1024)  // static smem (reduction scratch) counts against the 48 KB default too
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem) != cudaSuccess || occ < 1)
      occ = 1;
    if (occ > 4) occ = 4;
    const int64_t groups = (n + rows_per_cta - 1) / rows_per_cta;
    const int64_t cap = static_cast<int64_t>(sm_count()) * occ;
    grid = static_cast<int>(groups < cap ? groups : cap);
    n_part = static_cast<int64_t>(grid) * rows_per_cta;
    kern<<<grid, kThreads, smem, st>>>(dxx, p0, p1, dyy, aa, s0, s1, ww, emb, n, offset);
  };
  if (plan.mode == 1) {
    if (nvec <= 32) run(rowwise_bwd_vec<T, 32, 1, OP>, 8, 1);
    else if (nvec <= 64) run(rowwise_bwd_vec<T, 32, 2, OP>, 8, 2);
    else run(rowwise_bwd_vec<T, 32, 4, OP>, 8, 4);
  } else if (plan.mode == 2) {
    if (nvec <= 256) run(rowwise_bwd_vec<T, 256, 1, OP>, 1, 1);
    else if (nvec <= 512) run(rowwise_bwd_vec<T, 256, 2, OP>, 1, 2);
    else if (nvec <= 1024) run(rowwise_bwd_vec<T, 256, 4, OP>, 1, 4);
    else if constexpr (OP == 0) run(rowwise_bwd_vec<T, 256, 8, OP>, 1, 8);
  } else {
    run(rowwise_bwd_generic<T, OP>, 1, 0);
  }
  NNOP_LAUNCH_CHECK();
  if (OP != 0) {
    // every (CTA, sub-row) writes its partial row (zeros if it never saw a live row)
    const int64_t live_part = n_part;
    const dim3 blk(32, 8);
    const unsigned grid = static_cast<unsigned>((emb + 31) / 32);
    reduce_partials<TW><<<grid, blk, 0, st>>>(dw, p0, live_part, emb);
    if (OP == 2) reduce_partials<TW><<<grid, blk, 0, st>>>(db, p1, live_part, emb);
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
