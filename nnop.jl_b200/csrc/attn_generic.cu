// attn_generic.cu -- SIMT flash attention (forward, dK/dV, dQ) for everything the tcgen05
// path does not take: Float32, E in {16,32,256}, `pair` bias / `dpair`, `kpad_mask`.
// Same maths as the reference kernels (src/attention.jl:49-121, src/attention_bwd.jl:39-160)
// but: fp32 accumulation and statistics for every T, one log-sum-exp residual instead of
// (ms, ls), no per-tile renormalisation of O, dK/dV reduced inside the CTA over the q-heads of
// a GQA group (no global atomics), dQ in its own kernel (no global read-modify-write), and
// coalesced flat tile loads instead of the reference's stride-E row walks (:30-35).
//
// Thread layout: a query row (or key row in the dK/dV kernel) is shared by TPR adjacent lanes;
// lane `si` owns the float4 chunks {si + TPR*c} of the E axis, so the TPR partial dot products
// of a (row, key) pair are combined with log2(TPR) shuffles and K/V (or Q/dO) tiles are read
// from shared memory as conflict-free broadcast float4s.
#include "common.cuh"
#include "internal.h"

namespace nnop {
namespace {

constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct GParams {
  void* o; float* lse;
  const void* q; const void* k; const void* v; const void* pair; const uint8_t* kpad;
  void* dq; void* dk; void* dv; void* dpair; const void* dO; const float* delta;
  int QL, KL, QH, KH, B, causal;
  float scale;
};

template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load4<__half>(const __half* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
  float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void store4<__half>(__half* p, float4 v) {
  uint2 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
  h[0] = __floats2half2_rn(v.x, v.y);
  h[1] = __floats2half2_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
  h[0] = __floats2bfloat162_rn(v.x, v.y);
  h[1] = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}

// owned slice of one global row -> registers (zeros when !valid)
template <typename T, int E, int TPR>
__device__ __forceinline__ void load_slice(const T* row, int si, bool valid, float (&out)[E / TPR]) {
  constexpr int NC = E / TPR / 4;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float4 v = valid ? load4<T>(row + 4 * (si + TPR * c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    out[4 * c] = v.x; out[4 * c + 1] = v.y; out[4 * c + 2] = v.z; out[4 * c + 3] = v.w;
  }
}
template <typename T, int E, int TPR>
__device__ __forceinline__ void store_slice(T* row, int si, const float (&in)[E / TPR], float mul) {
  constexpr int NC = E / TPR / 4;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    store4<T>(row + 4 * (si + TPR * c), make_float4(in[4 * c] * mul, in[4 * c + 1] * mul,
                                                    in[4 * c + 2] * mul, in[4 * c + 3] * mul));
}
// ROWS x E tile (contiguous rows in global) -> fp32 shared memory, zero-filled past rows_valid
template <typename T, int E, int ROWS>
__device__ __forceinline__ void load_tile(float* sm, const T* g, int rows_valid) {
  for (int i = threadIdx.x * 4; i < ROWS * E; i += kThreads * 4) {
    const int r = i / E;
    float4 v = (r < rows_valid) ? load4<T>(g + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(sm + i) = v;
  }
}
template <int E, int TPR>
__device__ __forceinline__ float dot_slice(const float (&a)[E / TPR], const float* sm_row, int si) {
  constexpr int NC = E / TPR / 4;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float4 kv = *reinterpret_cast<const float4*>(sm_row + 4 * (si + TPR * c));
    acc = fmaf(a[4 * c], kv.x, acc);
    acc = fmaf(a[4 * c + 1], kv.y, acc);
    acc = fmaf(a[4 * c + 2], kv.z, acc);
    acc = fmaf(a[4 * c + 3], kv.w, acc);
  }
  return acc;
}
template <int E, int TPR>
__device__ __forceinline__ void axpy_slice(float (&acc)[E / TPR], float a, const float* sm_row,
                                           int si) {
  constexpr int NC = E / TPR / 4;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(sm_row + 4 * (si + TPR * c));
    acc[4 * c] = fmaf(a, v.x, acc[4 * c]);
    acc[4 * c + 1] = fmaf(a, v.y, acc[4 * c + 1]);
    acc[4 * c + 2] = fmaf(a, v.z, acc[4 * c + 2]);
    acc[4 * c + 3] = fmaf(a, v.w, acc[4 * c + 3]);
  }
}
template <int TPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = 1; o < TPR; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
template <typename T, int E, int TPR, int BN>
__global__ void __launch_bounds__(kThreads) attn_fwd_generic_kernel(GParams p) {
  constexpr int EPT = E / TPR;
  constexpr int BM = kThreads / TPR;
  __shared__ __align__(16) float Ks[BN * E];
  __shared__ __align__(16) float Vs[BN * E];
  const int si = threadIdx.x % TPR;
  const int row = blockIdx.x * BM + threadIdx.x / TPR;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.QH / p.KH);
  const bool valid = row < p.QL;
  const T* qg = static_cast<const T*>(p.q) + ((static_cast<int64_t>(b) * p.QH + h) * p.QL + row) * E;
  const T* kg = static_cast<const T*>(p.k) + (static_cast<int64_t>(b) * p.KH + hk) * p.KL * E;
  const T* vg = static_cast<const T*>(p.v) + (static_cast<int64_t>(b) * p.KH + hk) * p.KL * E;
  const T* pair = static_cast<const T*>(p.pair);

  float q[EPT], o[EPT];
  load_slice<T, E, TPR>(qg, si, valid, q);
#pragma unroll
  for (int i = 0; i < EPT; ++i) o[i] = 0.f;
  float m = -INFINITY, l = 0.f;

  int kmax = p.KL;
  if (p.causal) kmax = min(p.KL, (static_cast<int>(blockIdx.x) + 1) * BM);
  const float sl2 = p.scale * kLog2e;

  for (int k0 = 0; k0 < kmax; k0 += BN) {
    __syncthreads();
    load_tile<T, E, BN>(Ks, kg + static_cast<int64_t>(k0) * E, p.KL - k0);
    load_tile<T, E, BN>(Vs, vg + static_cast<int64_t>(k0) * E, p.KL - k0);
    __syncthreads();
    float s[BN];
    float tmax = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < BN; ++kk) {
      float d = group_sum<TPR>(dot_slice<E, TPR>(q, Ks + kk * E, si));
      const int kidx = k0 + kk;
      d *= sl2;  // logits kept in log2 units
      bool keep = kidx < p.KL && valid;
      if (p.causal) keep = keep && kidx <= row;
      if (keep && p.kpad) keep = p.kpad[static_cast<int64_t>(b) * p.KL + kidx] != 0;
      if (keep && pair)
        d = fmaf(to_f32<T>(pair[((static_cast<int64_t>(b) * p.KL + kidx) * p.QL + row) * p.QH + h]),
                 kLog2e, d);
      s[kk] = keep ? d : -INFINITY;
      tmax = fmaxf(tmax, s[kk]);
    }
    const float m_new = fmaxf(m, tmax);
    if (m_new > -INFINITY) {
      const float alpha = fast_exp2(m - m_new);  // m == -inf -> 0
      float psum = 0.f;
#pragma unroll
      for (int i = 0; i < EPT; ++i) o[i] *= alpha;
#pragma unroll
      for (int kk = 0; kk < BN; ++kk) {
        const float pv = fast_exp2(s[kk] - m_new);
        psum += pv;
        axpy_slice<E, TPR>(o, pv, Vs + kk * E, si);
      }
      l = fmaf(l, alpha, psum);
      m = m_new;
    }
  }
  if (valid) {
    const float inv = l > 0.f ? 1.f / l : 0.f;
    T* og = static_cast<T*>(p.o) + ((static_cast<int64_t>(b) * p.QH + h) * p.QL + row) * E;
    store_slice<T, E, TPR>(og, si, o, inv);
    if (si == 0)
      p.lse[(static_cast<int64_t>(b) * p.QH + h) * p.QL + row] =
          l > 0.f ? (m + fast_log2(l)) * 0.6931471805599453f : -INFINITY;
  }
}

// ---------------------------------------------------------------------------------------
// backward: dK, dV.  CTA = BMK keys of one kv head; loops over the group's q heads and over
// q tiles of BN rows staged in shared memory.
// ---------------------------------------------------------------------------------------
template <typename T, int E, int TPR, int BN>
__global__ void __launch_bounds__(kThreads) attn_bwd_dkdv_generic_kernel(GParams p) {
  constexpr int EPT = E / TPR;
  constexpr int BMK = kThreads / TPR;
  __shared__ __align__(16) float Qs[BN * E];
  __shared__ __align__(16) float Ds[BN * E];
  __shared__ float lse_s[BN];
  __shared__ float del_s[BN];
  const int si = threadIdx.x % TPR;
  const int key = blockIdx.x * BMK + threadIdx.x / TPR;
  const int hk = blockIdx.y, b = blockIdx.z;
  const int g = p.QH / p.KH;
  const bool kvalid = key < p.KL;
  const int64_t kvoff = ((static_cast<int64_t>(b) * p.KH + hk) * p.KL + key) * E;
  const T* pair = static_cast<const T*>(p.pair);
  T* dpair = static_cast<T*>(p.dpair);

  float kr[EPT], vr[EPT], dk[EPT], dv[EPT];
  load_slice<T, E, TPR>(static_cast<const T*>(p.k) + kvoff, si, kvalid, kr);
  load_slice<T, E, TPR>(static_cast<const T*>(p.v) + kvoff, si, kvalid, vr);
#pragma unroll
  for (int i = 0; i < EPT; ++i) dk[i] = dv[i] = 0.f;
  bool kkeep = kvalid;
  if (kkeep && p.kpad) kkeep = p.kpad[static_cast<int64_t>(b) * p.KL + key] != 0;
  const float sl2 = p.scale * kLog2e;

  // causal: q rows below the CTA's first key see none of its keys (skip unless dpair needs 0s)
  int q_begin = 0;
  if (p.causal && !dpair) q_begin = (blockIdx.x * BMK / BN) * BN;

  for (int hq = hk * g; hq < (hk + 1) * g; ++hq) {
    const int64_t qbase = (static_cast<int64_t>(b) * p.QH + hq) * p.QL;
    for (int q0 = q_begin; q0 < p.QL; q0 += BN) {
      __syncthreads();
      load_tile<T, E, BN>(Qs, static_cast<const T*>(p.q) + (qbase + q0) * E, p.QL - q0);
      load_tile<T, E, BN>(Ds, static_cast<const T*>(p.dO) + (qbase + q0) * E, p.QL - q0);
      if (threadIdx.x < BN) {
        const bool rv = q0 + threadIdx.x < p.QL;
        lse_s[threadIdx.x] = rv ? p.lse[qbase + q0 + threadIdx.x] * kLog2e : -INFINITY;
        del_s[threadIdx.x] = rv ? p.delta[qbase + q0 + threadIdx.x] : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int r = 0; r < BN; ++r) {
        const int qrow = q0 + r;
        float s = group_sum<TPR>(dot_slice<E, TPR>(kr, Qs + r * E, si)) * sl2;
        const float dp = group_sum<TPR>(dot_slice<E, TPR>(vr, Ds + r * E, si));
        bool keep = kkeep && qrow < p.QL;
        if (p.causal) keep = keep && key <= qrow;
        const int64_t pidx = ((static_cast<int64_t>(b) * p.KL + key) * p.QL + qrow) * p.QH + hq;
        if (keep && pair) s = fmaf(to_f32<T>(pair[pidx]), kLog2e, s);
        const float ls = lse_s[r];
        const float pv = (keep && ls > -INFINITY) ? fast_exp2(s - ls) : 0.f;
        const float ds = pv * (dp - del_s[r]);
        axpy_slice<E, TPR>(dv, pv, Ds + r * E, si);
        axpy_slice<E, TPR>(dk, ds, Qs + r * E, si);
        if (dpair && si == 0 && kvalid && qrow < p.QL) dpair[pidx] = from_f32<T>(ds);
      }
    }
  }
  if (kvalid) {
    store_slice<T, E, TPR>(static_cast<T*>(p.dk) + kvoff, si, dk, p.scale);
    store_slice<T, E, TPR>(static_cast<T*>(p.dv) + kvoff, si, dv, 1.f);
  }
}

// ---------------------------------------------------------------------------------------
// backward: dQ.  Same structure as the forward.
// ---------------------------------------------------------------------------------------
template <typename T, int E, int TPR, int BN>
__global__ void __launch_bounds__(kThreads) attn_bwd_dq_generic_kernel(GParams p) {
  constexpr int EPT = E / TPR;
  constexpr int BM = kThreads / TPR;
  __shared__ __align__(16) float Ks[BN * E];
  __shared__ __align__(16) float Vs[BN * E];
  const int si = threadIdx.x % TPR;
  const int row = blockIdx.x * BM + threadIdx.x / TPR;
  const int h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (p.QH / p.KH);
  const bool valid = row < p.QL;
  const int64_t qoff = ((static_cast<int64_t>(b) * p.QH + h) * p.QL + row) * E;
  const T* kg = static_cast<const T*>(p.k) + (static_cast<int64_t>(b) * p.KH + hk) * p.KL * E;
  const T* vg = static_cast<const T*>(p.v) + (static_cast<int64_t>(b) * p.KH + hk) * p.KL * E;
  const T* pair = static_cast<const T*>(p.pair);

  float q[EPT], dO[EPT], dq[EPT];
  load_slice<T, E, TPR>(static_cast<const T*>(p.q) + qoff, si, valid, q);
  load_slice<T, E, TPR>(static_cast<const T*>(p.dO) + qoff, si, valid, dO);
#pragma unroll
  for (int i = 0; i < EPT; ++i) dq[i] = 0.f;
  const int64_t sidx = (static_cast<int64_t>(b) * p.QH + h) * p.QL + row;
  const float ls = valid ? p.lse[sidx] * kLog2e : -INFINITY;
  const float del = valid ? p.delta[sidx] : 0.f;
  const float sl2 = p.scale * kLog2e;

  int kmax = p.KL;
  if (p.causal) kmax = min(p.KL, (static_cast<int>(blockIdx.x) + 1) * BM);
  for (int k0 = 0; k0 < kmax; k0 += BN) {
    __syncthreads();
    load_tile<T, E, BN>(Ks, kg + static_cast<int64_t>(k0) * E, p.KL - k0);
    load_tile<T, E, BN>(Vs, vg + static_cast<int64_t>(k0) * E, p.KL - k0);
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < BN; ++kk) {
      const int kidx = k0 + kk;
      float s = group_sum<TPR>(dot_slice<E, TPR>(q, Ks + kk * E, si)) * sl2;
      const float dp = group_sum<TPR>(dot_slice<E, TPR>(dO, Vs + kk * E, si));
      bool keep = kidx < p.KL && valid;
      if (p.causal) keep = keep && kidx <= row;
      if (keep && p.kpad) keep = p.kpad[static_cast<int64_t>(b) * p.KL + kidx] != 0;
      if (keep && pair)
        s = fmaf(to_f32<T>(pair[((static_cast<int64_t>(b) * p.KL + kidx) * p.QL + row) * p.QH + h]),
                 kLog2e, s);
      const float pv = (keep && ls > -INFINITY) ? fast_exp2(s - ls) : 0.f;
      const float ds = pv * (dp - del);
      axpy_slice<E, TPR>(dq, ds, Ks + kk * E, si);
    }
  }
  if (valid) store_slice<T, E, TPR>(static_cast<T*>(p.dq) + qoff, si, dq, p.scale);
}

// ---------------------------------------------------------------------------------------
// delta[b,h,q] = sum_e dO[b,h,q,e] * O[b,h,q,e]      (src/attention_bwd.jl:190-196; the
// reference's dO/l pre-scaling is not needed because O is stored normalised here)
// one 16-byte vector of dO and of O per lane, LPR lanes per row.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
attn_bwd_delta_kernel(float* __restrict__ delta, const T* __restrict__ dO, const T* __restrict__ o,
                      int64_t rows, int E) {
  constexpr int VE = 16 / sizeof(T);
  const int nv = E / VE;                 // vectors per row (E >= 16 -> nv >= 2)
  const int lpr = nv < 32 ? nv : 32;     // lanes per row (power of two)
  const int vpl = nv / lpr;              // vectors per lane
  const int rows_per_warp = 32 / lpr;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
  const int64_t row = warp * rows_per_warp + lane / lpr;
  const int li = lane % lpr;
  float acc = 0.f;
  if (row < rows) {
    for (int i = 0; i < vpl; ++i) {
      const int64_t off = row * E + static_cast<int64_t>(li + i * lpr) * VE;
      if constexpr (sizeof(T) == 4) {
        const float4 a = *reinterpret_cast<const float4*>(dO + off);
        const float4 c = *reinterpret_cast<const float4*>(o + off);
        acc += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
      } else {
        const uint4 a = *reinterpret_cast<const uint4*>(dO + off);
        const uint4 c = *reinterpret_cast<const uint4*>(o + off);
        const T* ah = reinterpret_cast<const T*>(&a);
        const T* ch = reinterpret_cast<const T*>(&c);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(to_f32<T>(ah[j]), to_f32<T>(ch[j]), acc);
      }
    }
  }
  for (int off = 1; off < lpr; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (row < rows && li == 0) delta[row] = acc;
}

GParams to_g(const AttnParams& a) {
  GParams g;
  g.o = a.o; g.lse = a.lse; g.q = a.q; g.k = a.k; g.v = a.v; g.pair = a.pair; g.kpad = a.kpad;
  g.dq = a.dq; g.dk = a.dk; g.dv = a.dv; g.dpair = a.dpair; g.dO = a.dO; g.delta = a.delta;
  g.QL = a.QL; g.KL = a.KL; g.QH = a.QH; g.KH = a.KH; g.B = a.B; g.causal = a.causal;
  g.scale = a.scale;
  return g;
}

template <typename T, int E, int TPR, int BN>
int launch_generic(const AttnParams& a, int which) {
  const GParams g = to_g(a);
  constexpr int BM = kThreads / TPR;
  if (which == 0) {
    dim3 grid((a.QL + BM - 1) / BM, a.QH, a.B);
    attn_fwd_generic_kernel<T, E, TPR, BN><<<grid, kThreads, 0, a.stream>>>(g);
  } else {
    dim3 gk((a.KL + BM - 1) / BM, a.KH, a.B);
    attn_bwd_dkdv_generic_kernel<T, E, TPR, BN><<<gk, kThreads, 0, a.stream>>>(g);
    dim3 gq((a.QL + BM - 1) / BM, a.QH, a.B);
    attn_bwd_dq_generic_kernel<T, E, TPR, BN><<<gq, kThreads, 0, a.stream>>>(g);
  }
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

template <typename T>
int dispatch_E(const AttnParams& a, int which) {
  switch (a.E) {
    case 16: return launch_generic<T, 16, 4, 32>(a, which);
    case 32: return launch_generic<T, 32, 4, 32>(a, which);
    case 64: return launch_generic<T, 64, 4, 32>(a, which);
    case 128: return launch_generic<T, 128, 8, 32>(a, which);
    case 256: return launch_generic<T, 256, 8, 16>(a, which);
    default:
      return fail(NNOP_ERR_UNSUPPORTED_E,
                  "Embedding dim `%d` is not supported (power of 2 in [16, 256]).", a.E);
  }
}

int dispatch(const AttnParams& a, int which) {
  switch (a.dtype) {
    case NNOP_F32: return dispatch_E<float>(a, which);
    case NNOP_F16: return dispatch_E<__half>(a, which);
    case NNOP_BF16: return dispatch_E<__nv_bfloat16>(a, which);
    default: return fail(NNOP_ERR_DTYPE, "unknown dtype code %d", a.dtype);
  }
}

}  // namespace

int attn_generic_fwd(const AttnParams& a) { return dispatch(a, 0); }
int attn_generic_bwd(const AttnParams& a) { return dispatch(a, 1); }

int attn_bwd_preprocess(const AttnParams& a) {
  const int64_t rows = static_cast<int64_t>(a.B) * a.QH * a.QL;
  if (rows == 0) return NNOP_OK;
  const int ve = a.dtype == NNOP_F32 ? 4 : 8;
  const int nv = a.E / ve;
  const int lpr = nv < 32 ? nv : 32;
  const int rows_per_warp = 32 / lpr;
  const int64_t warps = (rows + rows_per_warp - 1) / rows_per_warp;
  const unsigned grid = static_cast<unsigned>((warps + kThreads / 32 - 1) / (kThreads / 32));
  switch (a.dtype) {
    case NNOP_F32:
      attn_bwd_delta_kernel<float><<<grid, kThreads, 0, a.stream>>>(
          a.delta, static_cast<const float*>(a.dO), static_cast<const float*>(a.o), rows, a.E);
      break;
    case NNOP_F16:
      attn_bwd_delta_kernel<__half><<<grid, kThreads, 0, a.stream>>>(
          a.delta, static_cast<const __half*>(a.dO), static_cast<const __half*>(a.o), rows, a.E);
      break;
    default:
      attn_bwd_delta_kernel<__nv_bfloat16><<<grid, kThreads, 0, a.stream>>>(
          a.delta, static_cast<const __nv_bfloat16*>(a.dO),
          static_cast<const __nv_bfloat16*>(a.o), rows, a.E);
  }
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}

}  // namespace nnop
