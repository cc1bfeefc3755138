// ring_ops.cu -- HBM-bound helpers of the sequence-sharded ("ring") attention variant
// (BASELINE config 5; additive -- the reference has no multi-GPU path):
//   nnop_attn_merge          fold one partial attention result (o_part, lse_part) of a K/V block
//                            into the running fp32 (o_acc, lse_acc) by log-sum-exp weights
//   nnop_accumulate_f32      acc (+)= T partial gradient (dq over steps, travelling dk/dv)
//   nnop_store_rows_from_f32 T(acc) into a row window of a (slabs, rows, E) output
// 128-bit accesses, one thread per 8 (16-bit) or 4 (fp32) elements.
#include "common.cuh"
#include <initializer_list>

#include "internal.h"

namespace nnop {
namespace {

template <typename T>
struct Vec {
  static constexpr int N = sizeof(T) == 4 ? 4 : 8;
};

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float (&x)[Vec<T>::N]) {
  if constexpr (sizeof(T) == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  } else {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = to_f32<T>(h[i]);
  }
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float (&x)[Vec<T>::N]) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
  } else {
    uint4 v;
    T* h = reinterpret_cast<T*>(&v);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = from_f32<T>(x[i]);
    *reinterpret_cast<uint4*>(p) = v;
  }
}

// one thread per vector; vpr = vectors per row.  lse_acc is updated by the row's first vector
// AFTER every vector of the row has read it: rows never straddle a CTA's read/write hazard
// because the new lse goes to a separate output array (lse_out may alias only when vpr == 1).
template <typename T>
__global__ void __launch_bounds__(256)
attn_merge_kernel(float* __restrict__ o_acc, const float* __restrict__ lse_acc,
                  float* __restrict__ lse_new, const T* __restrict__ o_part,
                  const float* __restrict__ lse_part, int vpr, int64_t nvec, int init) {
  constexpr int N = Vec<T>::N;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= nvec) return;
  const int64_t row = i / vpr;
  float xp[N];
  load_vec<T>(o_part + i * N, xp);
  const float lp = lse_part[row];
  float out[N];
  float ln;
  if (init) {
#pragma unroll
    for (int e = 0; e < N; ++e) out[e] = xp[e];
    ln = lp;
  } else {
    const float la = lse_acc[row];
    const float m = fmaxf(la, lp);
    if (m == -INFINITY) {  // both sides fully masked: stay at 0 / -inf
#pragma unroll
      for (int e = 0; e < N; ++e) out[e] = 0.f;
      ln = -INFINITY;
    } else {
      const float wa = __expf(la - m), wp = __expf(lp - m);
      const float inv = 1.f / (wa + wp);
      const float ca = wa * inv, cp = wp * inv;
      float xa[N];
#pragma unroll
      for (int e = 0; e < N; e += 4) {
        const float4 v = *reinterpret_cast<const float4*>(o_acc + i * N + e);
        xa[e] = v.x; xa[e + 1] = v.y; xa[e + 2] = v.z; xa[e + 3] = v.w;
      }
#pragma unroll
      for (int e = 0; e < N; ++e) out[e] = xa[e] * ca + xp[e] * cp;
      ln = m + __logf(wa + wp);
    }
  }
#pragma unroll
  for (int e = 0; e < N; e += 4)
    *reinterpret_cast<float4*>(o_acc + i * N + e) = make_float4(out[e], out[e + 1], out[e + 2], out[e + 3]);
  if (i % vpr == 0) lse_new[row] = ln;
}

template <typename T>
__global__ void __launch_bounds__(256)
accumulate_kernel(float* __restrict__ acc, const T* __restrict__ part, int64_t nvec, int init) {
  constexpr int N = Vec<T>::N;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= nvec) return;
  float x[N];
  load_vec<T>(part + i * N, x);
#pragma unroll
  for (int e = 0; e < N; e += 4) {
    float4 v = make_float4(x[e], x[e + 1], x[e + 2], x[e + 3]);
    if (!init) {
      const float4 a = *reinterpret_cast<const float4*>(acc + i * N + e);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    *reinterpret_cast<float4*>(acc + i * N + e) = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
store_rows_kernel(T* __restrict__ out, const float* __restrict__ acc, int vpr, int64_t rows,
                  int64_t out_slab_rows, int64_t out_row_offset, int64_t nvec) {
  constexpr int N = Vec<T>::N;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= nvec) return;
  const int64_t row = i / vpr, col = i % vpr;
  const int64_t slab = row / rows, r = row % rows;
  float x[N];
#pragma unroll
  for (int e = 0; e < N; e += 4) {
    const float4 v = *reinterpret_cast<const float4*>(acc + i * N + e);
    x[e] = v.x; x[e + 1] = v.y; x[e + 2] = v.z; x[e + 3] = v.w;
  }
  store_vec<T>(out + ((slab * out_slab_rows + out_row_offset + r) * vpr + col) * N, x);
}

inline unsigned blocks(int64_t n) { return static_cast<unsigned>((n + 255) / 256); }

// every kernel here moves float4 / uint4 vectors: all pointers must be 16-byte aligned
inline bool aligned16(std::initializer_list<const void*> ps) {
  uintptr_t x = 0;
  for (const void* p : ps) x |= reinterpret_cast<uintptr_t>(p);
  return (x & 15) == 0;
}

template <typename F>
int by_dtype(int dtype, F&& f) {
  if (dtype == NNOP_F32) return f(float{});
  if (dtype == NNOP_F16) return f(__half{});
  if (dtype == NNOP_BF16) return f(__nv_bfloat16{});
  return fail(NNOP_ERR_DTYPE, "unknown dtype code %d", dtype);
}

}  // namespace
}  // namespace nnop

using namespace nnop;

extern "C" int nnop_attn_merge(float* o_acc, float* lse_acc, float* lse_out, const void* o_part,
                               const float* lse_part, int dtype, int E, int64_t rows, int init,
                               void* stream) {
  clear_error();
  if (rows == 0) return NNOP_OK;
  if (!o_acc || !lse_acc || !lse_out || !o_part || !lse_part) return fail(NNOP_ERR_ARG, "NULL pointer");
  if (E <= 0 || rows < 0 || E % 8 != 0) return fail(NNOP_ERR_SHAPE, "E must be a positive multiple of 8");
  if (!aligned16({o_acc, o_part})) return fail(NNOP_ERR_ARG, "o_acc and o_part must be 16-byte aligned");
  if (lse_out == lse_acc && !init)
    return fail(NNOP_ERR_ARG, "lse_out must not alias lse_acc (rows are updated by several threads)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return by_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    const int vpr = E / Vec<T>::N;
    const int64_t nvec = rows * vpr;
    attn_merge_kernel<T><<<blocks(nvec), 256, 0, st>>>(o_acc, lse_acc, lse_out,
                                                     static_cast<const T*>(o_part), lse_part, vpr,
                                                     nvec, init);
    NNOP_LAUNCH_CHECK();
    return NNOP_OK;
  });
}

extern "C" int nnop_accumulate_f32(float* acc, const void* part, int dtype, int64_t n, int init,
                                   void* stream) {
  clear_error();
  if (n == 0) return NNOP_OK;
  if (!acc || !part) return fail(NNOP_ERR_ARG, "NULL pointer");
  const int vec = dtype == NNOP_F32 ? 4 : 8;  // elements per 128-bit vector of `part`
  if (n < 0 || n % vec != 0)
    return fail(NNOP_ERR_SHAPE, "element count must be a multiple of %d for this dtype", vec);
  if (!aligned16({acc, part})) return fail(NNOP_ERR_ARG, "acc and part must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return by_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    const int64_t nvec = n / Vec<T>::N;
    accumulate_kernel<T><<<blocks(nvec), 256, 0, st>>>(acc, static_cast<const T*>(part), nvec, init);
    NNOP_LAUNCH_CHECK();
    return NNOP_OK;
  });
}

extern "C" int nnop_store_rows_from_f32(void* out, const float* acc, int dtype, int E,
                                        int64_t n_slabs, int64_t rows, int64_t out_slab_rows,
                                        int64_t out_row_offset, void* stream) {
  clear_error();
  if (n_slabs * rows == 0) return NNOP_OK;
  if (!out || !acc) return fail(NNOP_ERR_ARG, "NULL pointer");
  if (E <= 0 || E % 8 != 0 || rows < 0 || out_row_offset < 0 || out_row_offset + rows > out_slab_rows)
    return fail(NNOP_ERR_SHAPE, "bad row window: rows=%lld offset=%lld slab rows=%lld",
                static_cast<long long>(rows), static_cast<long long>(out_row_offset),
                static_cast<long long>(out_slab_rows));
  if (!aligned16({out, acc})) return fail(NNOP_ERR_ARG, "out and acc must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return by_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    const int vpr = E / Vec<T>::N;
    const int64_t nvec = n_slabs * rows * vpr;
    store_rows_kernel<T><<<blocks(nvec), 256, 0, st>>>(static_cast<T*>(out), acc, vpr, rows,
                                                     out_slab_rows, out_row_offset, nvec);
    NNOP_LAUNCH_CHECK();
    return NNOP_OK;
  });
}
