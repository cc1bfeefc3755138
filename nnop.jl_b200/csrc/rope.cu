// rope.cu -- Llama rotate-half RoPE on q (E,L,QH,B) and k (E,L,KH,B) in one launch.
// Replaces llama_rope! (src/rope/llama_rope.jl:24-65) and the copy(q)/copy(k) before it
// (:75-76): out-of-place, so q and k are read once and written once.  The reference maps a
// thread to a sequence position and walks E with stride-E accesses; here consecutive threads
// walk the contiguous E axis with 128-bit loads (one vector from each half of the row).
//   out[i]       = x[i]*c - x[i+E/2]*s
//   out[i+E/2]   = x[i+E/2]*c + x[i]*s        c = cos[i,l,b], s = sin[i,l,b]*sin_sign  (:43-61)
// fp32 math for every T (the reference's T*Float32 promotion), HBM-bound.
#include "common.cuh"
#include "internal.h"

namespace nnop {
namespace {

template <typename T, int VE>
__device__ __forceinline__ void ld(const T* p, float (&o)[VE]) {
  if constexpr (VE == 1) {
    o[0] = to_f32<T>(*p);
  } else if constexpr (sizeof(T) == 4) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  } else {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    const T* h = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = to_f32<T>(h[i]);
  }
}
template <typename T, int VE>
__device__ __forceinline__ void st(T* p, const float (&o)[VE]) {
  if constexpr (VE == 1) {
    *p = from_f32<T>(o[0]);
  } else if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
    uint4 v;
    T* h = reinterpret_cast<T*>(&v);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = from_f32<T>(o[i]);
    *reinterpret_cast<uint4*>(p) = v;
  }
}
template <int VE>
__device__ __forceinline__ void ldf(const float* p, float (&o)[VE]) {
  if constexpr (VE == 1) {
    o[0] = *p;
  } else {
#pragma unroll
    for (int i = 0; i < VE; i += 4) {
      float4 v = *reinterpret_cast<const float4*>(p + i);
      o[i] = v.x; o[i + 1] = v.y; o[i + 2] = v.z; o[i + 3] = v.w;
    }
  }
}

template <typename T, int VE>
__global__ void __launch_bounds__(256)
llama_rope_kernel(T* q_out, T* k_out, const T* q_in, const T* k_in,
                  const float* __restrict__ cosp, const float* __restrict__ sinp, int E, int64_t L,
                  int QH, int KH, int B, float sin_sign) {
  const int half = E / 2;
  const int HV = half / VE;
  const int HT = QH + KH;
  const int64_t total = static_cast<int64_t>(B) * HT * L * HV;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(idx % HV);
    int64_t r = idx / HV;
    const int64_t l = r % L;
    r /= L;
    const int hh = static_cast<int>(r % HT);
    const int64_t b = r / HT;
    const T* src;
    T* dst;
    if (hh < QH) {
      const int64_t off = ((b * QH + hh) * L + l) * E;
      src = q_in + off;
      dst = q_out + off;
    } else {
      const int64_t off = ((b * KH + (hh - QH)) * L + l) * E;
      src = k_in + off;
      dst = k_out + off;
    }
    const int64_t coff = (b * L + l) * E + static_cast<int64_t>(j) * VE;
    float x1[VE], x2[VE], c[VE], s[VE], o1[VE], o2[VE];
    ld<T, VE>(src + j * VE, x1);
    ld<T, VE>(src + half + j * VE, x2);
    ldf<VE>(cosp + coff, c);
    ldf<VE>(sinp + coff, s);
#pragma unroll
    for (int i = 0; i < VE; ++i) {
      const float sv = s[i] * sin_sign;
      o1[i] = x1[i] * c[i] - x2[i] * sv;
      o2[i] = x2[i] * c[i] + x1[i] * sv;
    }
    st<T, VE>(dst + j * VE, o1);
    st<T, VE>(dst + half + j * VE, o2);
  }
}

template <typename T>
int launch_rope(void* q_out, void* k_out, const void* q_in, const void* k_in, const float* cosp,
                const float* sinp, int E, int64_t L, int QH, int KH, int B, float sin_sign,
                cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  const int half = E / 2;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = half % VEC == 0 && al(q_out) && al(k_out) && al(q_in) && al(k_in) && al(cosp) &&
                   al(sinp);
  const int ve = vec ? VEC : 1;
  const int64_t total = static_cast<int64_t>(B) * (QH + KH) * L * (half / ve);
  if (total == 0) return NNOP_OK;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 32;
  if (blocks > cap) blocks = cap;
  if (vec)
    llama_rope_kernel<T, VEC><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        static_cast<T*>(q_out), static_cast<T*>(k_out), static_cast<const T*>(q_in),
        static_cast<const T*>(k_in), cosp, sinp, E, L, QH, KH, B, sin_sign);
  else
    llama_rope_kernel<T, 1><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        static_cast<T*>(q_out), static_cast<T*>(k_out), static_cast<const T*>(q_in),
        static_cast<const T*>(k_in), cosp, sinp, E, L, QH, KH, B, sin_sign);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

}  // namespace
}  // namespace nnop

using namespace nnop;

extern "C" int nnop_llama_rope(void* q_out, void* k_out, const void* q_in, const void* k_in,
                               const float* cosp, const float* sinp, int dtype, int E, int64_t L,
                               int QH, int KH, int B, float sin_sign, void* stream) {
  clear_error();
  if (E <= 0 || (E & 1)) return fail(NNOP_ERR_SHAPE, "RoPE head dim `%d` must be even.", E);
  if (L < 0 || QH < 0 || KH < 0 || B < 0) return fail(NNOP_ERR_SHAPE, "negative size");
  if (static_cast<int64_t>(B) * (QH + KH) * L == 0) return NNOP_OK;
  if ((QH > 0 && (!q_out || !q_in)) || (KH > 0 && (!k_out || !k_in)) || !cosp || !sinp)
    return fail(NNOP_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (dtype) {
    case NNOP_F32:
      return launch_rope<float>(q_out, k_out, q_in, k_in, cosp, sinp, E, L, QH, KH, B, sin_sign, st);
    case NNOP_F16:
      return launch_rope<__half>(q_out, k_out, q_in, k_in, cosp, sinp, E, L, QH, KH, B, sin_sign, st);
    case NNOP_BF16:
      return launch_rope<__nv_bfloat16>(q_out, k_out, q_in, k_in, cosp, sinp, E, L, QH, KH, B,
                                        sin_sign, st);
    default:
      return fail(NNOP_ERR_DTYPE, "unknown dtype code %d", dtype);
  }
}
