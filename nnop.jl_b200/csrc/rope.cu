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

// packed fp32x2 arithmetic (two lanes per issue slot on sm_100)
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

// Thread mapping (r02; the r01 kernel spent 72 % of its issue slots, mostly on three 64-bit divisions per
// vector): a thread keeps ONE vector position j of the half row for its whole life and a CTA owns a block of
// sequence positions of ONE (batch, head) -- blockIdx.x = (b * HT + hh) * nlb + lb -- so the only divisions
// are two 32-bit ones per thread, once.  Rows are then walked with constant strides.
template <typename T, int VE>
__global__ void __launch_bounds__(256)
llama_rope_kernel(T* q_out, T* k_out, const T* q_in, const T* k_in,
                  const float* __restrict__ cosp, const float* __restrict__ sinp, int E, int64_t L,
                  int QH, int KH, int B, float sin_sign, int HV, int rpb, int nlb, int rows_per_cta) {
  const int half = E / 2;
  const int HT = QH + KH;
  const int j = threadIdx.x % HV;
  const int rl = threadIdx.x / HV;            // row of the CTA's current group of rpb rows
  if (rl >= rpb) return;
  const unsigned bh = blockIdx.x / nlb, lb = blockIdx.x - bh * nlb;
  const int b = bh / HT, hh = bh - b * HT;
  const bool is_q = hh < QH;
  const int64_t head_off = is_q ? (static_cast<int64_t>(b) * QH + hh) * L * E
                                : (static_cast<int64_t>(b) * KH + (hh - QH)) * L * E;
  const T* src = (is_q ? q_in : k_in) + head_off + j * VE;
  T* dst = (is_q ? q_out : k_out) + head_off + j * VE;
  const float* cp = cosp + static_cast<int64_t>(b) * L * E + j * VE;
  const float* sp = sinp + static_cast<int64_t>(b) * L * E + j * VE;
  const int64_t l0 = static_cast<int64_t>(lb) * rows_per_cta;
  const int64_t l1 = l0 + rows_per_cta < L ? l0 + rows_per_cta : L;
  for (int64_t l = l0 + rl; l < l1; l += rpb) {
    const int64_t ro = l * E;
    float x1[VE], x2[VE], c[VE], sn[VE], o1[VE], o2[VE];
    ld<T, VE>(src + ro, x1);
    ld<T, VE>(src + ro + half, x2);
    ldf<VE>(cp + ro, c);
    ldf<VE>(sp + ro, sn);
    if constexpr (VE % 2 == 0) {   // packed fp32x2: 6 operations per element pair instead of 10
      const float2 sg = make_float2(sin_sign, sin_sign), nsg = make_float2(-sin_sign, -sin_sign);
#pragma unroll
      for (int i = 0; i < VE; i += 2) {
        const float2 a = make_float2(x1[i], x1[i + 1]), bb = make_float2(x2[i], x2[i + 1]);
        const float2 cc = make_float2(c[i], c[i + 1]), ss = make_float2(sn[i], sn[i + 1]);
        const float2 sv = f2_mul(ss, sg), nsv = f2_mul(ss, nsg);
        const float2 r1 = f2_fma(bb, nsv, f2_mul(a, cc));     // x1 c - x2 s
        const float2 r2 = f2_fma(a, sv, f2_mul(bb, cc));      // x2 c + x1 s
        o1[i] = r1.x; o1[i + 1] = r1.y;
        o2[i] = r2.x; o2[i + 1] = r2.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < VE; ++i) {
        const float sv = sn[i] * sin_sign;
        o1[i] = x1[i] * c[i] - x2[i] * sv;
        o2[i] = x2[i] * c[i] + x1[i] * sv;
      }
    }
    st<T, VE>(dst + ro, o1);
    st<T, VE>(dst + ro + half, o2);
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
  const int HV = half / ve;
  if (HV > 256) return fail(NNOP_ERR_UNSUPPORTED_E, "RoPE head dim `%d` is not supported.", E);
  const int64_t total = static_cast<int64_t>(B) * (QH + KH) * L * HV;
  if (total == 0) return NNOP_OK;
  const int rpb = 256 / HV;                       // rows a CTA covers per pass
  // rows per CTA: enough passes to amortise the setup, enough CTAs (>= ~8 per SM) to fill the machine
  const int64_t bh = static_cast<int64_t>(B) * (QH + KH);
  int64_t passes = 4;   // measured: 2-4 passes best at config C3 (profiles/r02_perf_rope_passes.txt)
  while (passes > 1 && bh * ((L + rpb * passes - 1) / (rpb * passes)) < 24LL * sm_count()) passes >>= 1;
  const int rows_per_cta = static_cast<int>(rpb * passes);
  const int64_t nlb = (L + rows_per_cta - 1) / rows_per_cta;
  if (bh * nlb >= (1LL << 31)) return fail(NNOP_ERR_SHAPE, "RoPE problem too large for one launch");
  const unsigned blocks = static_cast<unsigned>(bh * nlb);
  if (vec)
    llama_rope_kernel<T, VEC><<<blocks, 256, 0, st>>>(
        static_cast<T*>(q_out), static_cast<T*>(k_out), static_cast<const T*>(q_in),
        static_cast<const T*>(k_in), cosp, sinp, E, L, QH, KH, B, sin_sign, HV, rpb, static_cast<int>(nlb),
        rows_per_cta);
  else
    llama_rope_kernel<T, 1><<<blocks, 256, 0, st>>>(
        static_cast<T*>(q_out), static_cast<T*>(k_out), static_cast<const T*>(q_in),
        static_cast<const T*>(k_in), cosp, sinp, E, L, QH, KH, B, sin_sign, HV, rpb, static_cast<int>(nlb),
        rows_per_cta);
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
