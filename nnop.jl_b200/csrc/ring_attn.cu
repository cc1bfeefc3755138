// ring_attn.cu -- sequence-sharded flash attention across the GPUs of one process (BASELINE config C5:
// L = 131072 on 8 x B200), behind the C ABI: nnop_ring_attn_fwd / nnop_ring_attn_bwd.
// Additive: the reference has no multi-GPU path; SURVEY.md 8(b) specifies the entry
// points (single process, per-device pointer arrays).
//
// Device d (rank d of W) owns a slice of the sequence axis of q, k, v (E, Ll, H, B).  Causal problems use
// the zig-zag layout -- the sequence is cut into 2W chunks of c rows and rank d owns chunks (d, 2W-1-d),
// concatenated -- which makes every step the same work on every rank and needs no mask beyond step 0:
//   step 0            the local problem: ONE causal call over the 2c local rows (chunk 2W-1-d lies after
//                     chunk d in the sequence, so top-left causal over the concatenation is exact)
//   step s, src < d   all 2c local queries x the FIRST chunk of rank src = (d-s) mod W, unmasked:
//                     only that chunk (half a block) crosses NVLink
//   step s, src > d   the local SECOND query chunk x both chunks of rank src, unmasked
// Every step is one dense nnop_flash_attn_fwd / _bwd call; partial (o, lse) are folded into fp32
// accumulators by log-sum-exp weights; the backward runs the same schedule on the FINAL o / lse, keeps dq
// local and sends each step's dk / dv partial (element type T, as the kernel writes it) to the block's
// owner, which adds it into its fp32 accumulator.
//
// Transport: K / V never hop rank to rank.  NVSwitch gives every GPU full bandwidth to every peer, so at
// step s rank d PULLS the block straight from its owner's (read-only) input tensors with
// cudaMemcpyPeerAsync / cudaMemcpy3DPeerAsync on a copy stream of its own, into a double-buffered landing
// area, one step ahead of the math; gradient partials are PUSHED to the owner the same way.  Copy engines
// move the data, so the attention grids (one CTA per SM, all shared memory) keep every SM.  Ordering is
// events only: the host thread never blocks, and every event is recorded (in host program order) before
// anything waits on it.  Streams and events made here are released before returning; nothing is retained.
#include <initializer_list>
#include <vector>

#include "common.cuh"
#include "internal.h"

namespace nnop {
namespace {

template <typename T>
struct Vec {
  static constexpr int N = sizeof(T) == 4 ? 4 : 8;
};
template <typename T>
__device__ __forceinline__ void ld_vec(const T* p, float (&x)[Vec<T>::N]) {
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
__device__ __forceinline__ void st_vec(T* p, const float (&x)[Vec<T>::N]) {
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

// Row addressing shared by the kernels below: `part` is (slabs, rows, E) dense; it addresses rows
// [acc_off, acc_off + rows) of every slab of `acc` (slabs, acc_rows, E).  vpr = vectors per row.
struct RowMap {
  int vpr;
  int64_t rows, acc_rows, acc_off, nvec;
  __device__ __forceinline__ int64_t acc_row(int64_t prow) const {
    return (prow / rows) * acc_rows + acc_off + prow % rows;
  }
};

// (o_acc, lse) <- log-sum-exp weighted combination with one partial result; lse updated in place: the
// vectors of a row sit in one CTA (256 % vpr == 0), all of them read lse before the barrier, one writes after
template <typename T>
__global__ void __launch_bounds__(256)
ring_merge_kernel(float* __restrict__ o_acc, float* __restrict__ lse, const T* __restrict__ o_part,
                  const float* __restrict__ lse_part, const RowMap m, int init) {
  constexpr int N = Vec<T>::N;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  const bool live = i < m.nvec;
  float out[N];
  float ln = 0.f;
  int64_t arow = 0, aoff = 0;
  if (live) {
    const int64_t prow = i / m.vpr;
    arow = m.acc_row(prow);
    aoff = (arow * m.vpr + i % m.vpr) * N;
    float xp[N];
    ld_vec<T>(o_part + i * N, xp);
    const float lp = lse_part[prow];
    if (init) {
#pragma unroll
      for (int e = 0; e < N; ++e) out[e] = xp[e];
      ln = lp;
    } else {
      const float la = lse[arow];
      const float mx = fmaxf(la, lp);
      if (mx == -INFINITY) {  // both sides fully masked: stay at 0 / -inf
#pragma unroll
        for (int e = 0; e < N; ++e) out[e] = 0.f;
        ln = -INFINITY;
      } else {
        const float wa = __expf(la - mx), wp = __expf(lp - mx);
        const float inv = 1.f / (wa + wp);
        const float ca = wa * inv, cp = wp * inv;
#pragma unroll
        for (int e = 0; e < N; e += 4) {
          const float4 a = *reinterpret_cast<const float4*>(o_acc + aoff + e);
          out[e] = a.x * ca + xp[e] * cp;
          out[e + 1] = a.y * ca + xp[e + 1] * cp;
          out[e + 2] = a.z * ca + xp[e + 2] * cp;
          out[e + 3] = a.w * ca + xp[e + 3] * cp;
        }
        ln = mx + __logf(wa + wp);
      }
    }
  }
  __syncthreads();
  if (live) {
#pragma unroll
    for (int e = 0; e < N; e += 4)
      *reinterpret_cast<float4*>(o_acc + aoff + e) = make_float4(out[e], out[e + 1], out[e + 2], out[e + 3]);
    if (i % m.vpr == 0) lse[arow] = ln;
  }
}

// acc rows (+)= float(part)
template <typename T>
__global__ void __launch_bounds__(256)
ring_accum_kernel(float* __restrict__ acc, const T* __restrict__ part, const RowMap m, int init) {
  constexpr int N = Vec<T>::N;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= m.nvec) return;
  const int64_t aoff = (m.acc_row(i / m.vpr) * m.vpr + i % m.vpr) * N;
  float x[N];
  ld_vec<T>(part + i * N, x);
#pragma unroll
  for (int e = 0; e < N; e += 4) {
    float4 v = make_float4(x[e], x[e + 1], x[e + 2], x[e + 3]);
    if (!init) {
      const float4 a = *reinterpret_cast<const float4*>(acc + aoff + e);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    *reinterpret_cast<float4*>(acc + aoff + e) = v;
  }
}

// out = T(acc), same dense layout
template <typename T>
__global__ void __launch_bounds__(256)
ring_store_kernel(T* __restrict__ out, const float* __restrict__ acc, int64_t nvec) {
  constexpr int N = Vec<T>::N;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= nvec) return;
  float x[N];
#pragma unroll
  for (int e = 0; e < N; e += 4) {
    const float4 v = *reinterpret_cast<const float4*>(acc + i * N + e);
    x[e] = v.x; x[e + 1] = v.y; x[e + 2] = v.z; x[e + 3] = v.w;
  }
  st_vec<T>(out + i * N, x);
}

inline unsigned nblk(int64_t n) { return static_cast<unsigned>((n + 255) / 256); }
inline size_t up256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

template <typename F>
int by_dtype(int dtype, F&& f) {
  if (dtype == NNOP_F32) return f(float{});
  if (dtype == NNOP_F16) return f(__half{});
  if (dtype == NNOP_BF16) return f(__nv_bfloat16{});
  return fail(NNOP_ERR_DTYPE, "unknown dtype code %d", dtype);
}

struct Dims {
  int dtype, E, Ll, QH, KH, B, W, causal;
  size_t es;        // element size
  int c;            // chunk rows (causal: Ll / 2, else Ll)
  size_t q_full() const { return static_cast<size_t>(B) * QH * Ll * E * es; }
  size_t kv_full() const { return static_cast<size_t>(B) * KH * Ll * E * es; }
  size_t q_rows() const { return static_cast<size_t>(B) * QH * Ll; }
  size_t kv_rows() const { return static_cast<size_t>(B) * KH * Ll; }
};

// per-device workspace layout (byte offsets); the same arithmetic serves the size queries
struct Layout {
  size_t kv[2];        // landing areas of a K / V block: [K | V], each up to kv_full
  size_t q2;           // chunk-contiguous copy of the second local q chunk (causal)
  size_t o_p, lse_p;   // forward: partial result of one step
  size_t o_acc;        // forward: fp32 accumulator (local layout)
  // backward
  size_t dO2, o2, lse2;   // chunk-contiguous copies (second chunk) of dO, o, lse
  size_t dq_p;            // dq of one step
  size_t part[2];         // [dk_p | dv_p] of one step, double-buffered (being pushed / being written)
  size_t recv[2];         // landing areas for other ranks' partials of MY block
  size_t dq_acc, dkv_acc; // fp32 accumulators
  size_t attn_ws, attn_ws_bytes;
  size_t total;
};

Layout make_layout(const Dims& d, bool backward) {
  Layout L{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t at = off;
    off += up256(bytes);
    return at;
  };
  L.kv[0] = take(2 * d.kv_full());
  L.kv[1] = take(2 * d.kv_full());
  L.q2 = take(d.causal ? d.q_full() / 2 : 0);
  if (!backward) {
    L.o_p = take(d.q_full());
    L.lse_p = take(d.q_rows() * sizeof(float));
    L.o_acc = take(d.q_rows() * d.E * sizeof(float));
  } else {
    L.dO2 = take(d.causal ? d.q_full() / 2 : 0);
    L.o2 = take(d.causal ? d.q_full() / 2 : 0);
    L.lse2 = take(d.causal ? d.q_rows() / 2 * sizeof(float) : 0);
    L.dq_p = take(d.q_full());
    for (int i = 0; i < 2; ++i) L.part[i] = take(2 * d.kv_full());
    for (int i = 0; i < 2; ++i) L.recv[i] = take(2 * d.kv_full());
    L.dq_acc = take(d.q_rows() * d.E * sizeof(float));
    L.dkv_acc = take(2 * d.kv_rows() * d.E * sizeof(float));
    size_t w = nnop_flash_attn_bwd_workspace_bytes(d.dtype, d.E, d.Ll, d.Ll, d.QH, d.KH, d.B);
    if (d.causal) {
      const size_t w1 = nnop_flash_attn_bwd_workspace_bytes(d.dtype, d.E, d.Ll, d.c, d.QH, d.KH, d.B);
      const size_t w2 = nnop_flash_attn_bwd_workspace_bytes(d.dtype, d.E, d.c, d.Ll, d.QH, d.KH, d.B);
      if (w1 > w) w = w1;
      if (w2 > w) w = w2;
    }
    L.attn_ws_bytes = w;
    L.attn_ws = take(w);
  }
  L.total = off;
  return L;
}

int check_dims(Dims& d, int dtype, int E, int Ll, int QH, int KH, int B, int ndev, int causal) {
  if (dtype != NNOP_F32 && dtype != NNOP_F16 && dtype != NNOP_BF16)
    return fail(NNOP_ERR_DTYPE, "unknown dtype code %d", dtype);
  if (ndev < 1 || ndev > 64) return fail(NNOP_ERR_ARG, "ring attention: ndev must be in [1, 64], got %d", ndev);
  if (E < 16 || E > 256 || (E & (E - 1)) != 0)
    return fail(NNOP_ERR_UNSUPPORTED_E, "Only power-of-2 embedding dims are supported.");
  if (Ll <= 0 || QH <= 0 || KH <= 0 || B <= 0)
    return fail(NNOP_ERR_SHAPE, "Invalid ring attention shape E=%d Ll=%d QH=%d KH=%d B=%d.", E, Ll, QH, KH, B);
  if (QH % KH != 0)
    return fail(NNOP_ERR_SHAPE, "Number of query heads `%d` must be divisible by number of KV heads `%d`.", QH, KH);
  if (causal && Ll % 2 != 0)
    return fail(NNOP_ERR_SHAPE, "causal ring attention needs an even local sequence length, got `%d`", Ll);
  d = Dims{dtype, E, Ll, QH, KH, B, ndev, causal ? 1 : 0, dtype_size(dtype), causal ? Ll / 2 : Ll};
  return NNOP_OK;
}

// Everything the two entry points share: per-rank copy streams, an event pool, the current-device guard.
struct Ring {
  const Dims& d;
  const int* devs;
  std::vector<cudaStream_t> comp, copy;
  std::vector<cudaEvent_t> pool;
  int saved_dev = 0;
  int rc = NNOP_OK;

  Ring(const Dims& dims, const int* devices, void* const* streams) : d(dims), devs(devices) {
    cudaGetDevice(&saved_dev);
    comp.resize(d.W);
    copy.assign(d.W, nullptr);
    for (int r = 0; r < d.W; ++r) comp[r] = static_cast<cudaStream_t>(streams ? streams[r] : nullptr);
  }
  ~Ring() {
    for (int r = 0; r < d.W; ++r)
      if (copy[r]) {
        cudaSetDevice(devs[r]);
        cudaStreamDestroy(copy[r]);   // pending work completes first (asynchronous release)
      }
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
    cudaSetDevice(saved_dev);
  }
  int on(int r) { return set(cudaSetDevice(devs[r])); }
  int set(cudaError_t e) {
    if (e != cudaSuccess && rc == NNOP_OK)
      rc = fail(NNOP_ERR_CUDA, "ring attention: CUDA call failed: %s", cudaGetErrorString(e));
    return rc;
  }
  int init() {
    for (int r = 0; r < d.W && rc == NNOP_OK; ++r) {
      on(r);
      set(cudaStreamCreateWithFlags(&copy[r], cudaStreamNonBlocking));
      for (int p = 0; p < d.W; ++p) {   // direct NVLink access to every peer (idempotent)
        if (devs[p] == devs[r]) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devs[r], devs[p]);
        if (can) {
          const cudaError_t e = cudaDeviceEnablePeerAccess(devs[p], 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) set(e);
          (void)cudaGetLastError();
        }
      }
    }
    return rc;
  }
  // new event on rank r's device, recorded on `st`
  cudaEvent_t record(int r, cudaStream_t st) {
    on(r);
    cudaEvent_t e = nullptr;
    set(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (e) {
      pool.push_back(e);
      set(cudaEventRecord(e, st));
    }
    return e;
  }
  void wait(cudaStream_t st, cudaEvent_t e) {
    if (e) set(cudaStreamWaitEvent(st, e, 0));
  }
  // rows [row0, row0 + rows) of every (b, h) slab of a (slabs, Ll, E) tensor on rank `from` -> dense
  // (slabs, rows, E) on rank `to`, enqueued on `st`
  void copy_rows(void* dst, int to, const void* src, int from, size_t slabs, int row0, int rows, cudaStream_t st) {
    const size_t row_bytes = static_cast<size_t>(d.E) * d.es;
    const char* s = static_cast<const char*>(src) + static_cast<size_t>(row0) * row_bytes;
    if (rows == d.Ll) {
      set(cudaMemcpyPeerAsync(dst, devs[to], s, devs[from], slabs * rows * row_bytes, st));
      return;
    }
    cudaMemcpy3DPeerParms p{};
    p.srcPtr = make_cudaPitchedPtr(const_cast<char*>(s), static_cast<size_t>(d.Ll) * row_bytes, rows * row_bytes, slabs);
    p.dstPtr = make_cudaPitchedPtr(dst, static_cast<size_t>(rows) * row_bytes, rows * row_bytes, slabs);
    p.extent = make_cudaExtent(static_cast<size_t>(rows) * row_bytes, slabs, 1);
    p.srcDevice = devs[from];
    p.dstDevice = devs[to];
    set(cudaMemcpy3DPeerAsync(&p, st));
  }
};

// what rank r computes at step s (s >= 1 when src != r): which query rows against which key rows
struct StepShape {
  int src;
  bool q_second;   // queries = second local chunk only (else all local rows)
  int QL, KL;      // rows of the call
  int causal;
};
StepShape step_shape(const Dims& d, int r, int s) {
  StepShape t;
  t.src = ((r - s) % d.W + d.W) % d.W;
  t.q_second = false;
  t.QL = t.KL = d.Ll;
  t.causal = 0;
  if (!d.causal) return t;
  if (s == 0) {
    t.causal = 1;
  } else if (t.src < r) {
    t.KL = d.c;            // first chunk of the block only
  } else {
    t.q_second = true;
    t.QL = d.c;
  }
  return t;
}

template <typename P>
int check_ptrs(P* const* arr, int n, const char* what) {
  if (!arr) return fail(NNOP_ERR_ARG, "ring attention: NULL pointer array `%s`", what);
  for (int i = 0; i < n; ++i)
    if (!arr[i] || (reinterpret_cast<uintptr_t>(arr[i]) & 15) != 0)
      return fail(NNOP_ERR_ARG, "ring attention: `%s[%d]` is NULL or not 16-byte aligned", what, i);
  return NNOP_OK;
}

}  // namespace
}  // namespace nnop

using namespace nnop;

extern "C" size_t nnop_ring_attn_fwd_workspace_bytes(int dtype, int E, int Ll, int QH, int KH, int B, int ndev,
                                                     int causal) {
  Dims d;
  if (check_dims(d, dtype, E, Ll, QH, KH, B, ndev, causal)) return 0;
  return make_layout(d, false).total;
}

extern "C" size_t nnop_ring_attn_bwd_workspace_bytes(int dtype, int E, int Ll, int QH, int KH, int B, int ndev,
                                                     int causal) {
  Dims d;
  if (check_dims(d, dtype, E, Ll, QH, KH, B, ndev, causal)) return 0;
  return make_layout(d, true).total;
}

extern "C" int nnop_ring_attn_fwd(void* const* o, float* const* lse, const void* const* q, const void* const* k,
                                  const void* const* v, const int* devices, int ndev, int dtype, int E, int Ll,
                                  int QH, int KH, int B, int causal, float scale, void* const* workspace,
                                  size_t workspace_bytes, void* const* streams) {
  clear_error();
  Dims d;
  if (int rc = check_dims(d, dtype, E, Ll, QH, KH, B, ndev, causal)) return rc;
  if (!devices) return fail(NNOP_ERR_ARG, "ring attention: NULL device list");
  if (int rc = check_ptrs(o, ndev, "o")) return rc;
  if (int rc = check_ptrs(lse, ndev, "lse")) return rc;
  if (int rc = check_ptrs(q, ndev, "q")) return rc;
  if (int rc = check_ptrs(k, ndev, "k")) return rc;
  if (int rc = check_ptrs(v, ndev, "v")) return rc;
  const Layout L = make_layout(d, false);
  if (workspace_bytes < L.total)
    return fail(NNOP_ERR_WORKSPACE, "ring attention forward needs a %zu-byte workspace per device, got %zu", L.total,
                workspace_bytes);
  if (!workspace) return fail(NNOP_ERR_WORKSPACE, "ring attention: NULL workspace array");
  for (int r = 0; r < ndev; ++r)
    if (!workspace[r] || (reinterpret_cast<uintptr_t>(workspace[r]) & 255) != 0)
      return fail(NNOP_ERR_WORKSPACE, "workspace[%d] must be non-NULL and 256-byte aligned", r);

  Ring R(d, devices, streams);
  if (R.init()) return R.rc;
  const int W = d.W;
  auto ws = [&](int r, size_t off) { return static_cast<char*>(workspace[r]) + off; };
  const size_t kvb = d.kv_full();
  const size_t kv_slabs = static_cast<size_t>(B) * KH, q_slabs = static_cast<size_t>(B) * QH;

  std::vector<cudaEvent_t> in_ready(W), kv_ready(W, nullptr);
  std::vector<std::vector<cudaEvent_t>> readers(W);   // pulls that read rank r's k / v (other ranks' copy streams)
  std::vector<std::vector<cudaEvent_t>> comp_done(W, std::vector<cudaEvent_t>(W, nullptr));
  for (int r = 0; r < W; ++r) in_ready[r] = R.record(r, R.comp[r]);
  if (d.causal)
    for (int r = 0; r < W; ++r) {   // second q chunk, chunk-contiguous
      R.on(r);
      R.copy_rows(ws(r, L.q2), r, q[r], r, q_slabs, d.c, d.c, R.comp[r]);
    }

  auto pull = [&](int r, int s) {   // K / V block of step s into landing area s & 1, on rank r's copy stream
    const StepShape t = step_shape(d, r, s);
    R.on(r);
    R.wait(R.copy[r], in_ready[r]);
    R.wait(R.copy[r], in_ready[t.src]);
    if (s >= 3) R.wait(R.copy[r], comp_done[r][s - 2]);
    char* dst = ws(r, L.kv[s & 1]);
    R.copy_rows(dst, r, k[t.src], t.src, kv_slabs, 0, t.KL, R.copy[r]);
    R.copy_rows(dst + kvb, r, v[t.src], t.src, kv_slabs, 0, t.KL, R.copy[r]);
    kv_ready[r] = R.record(r, R.copy[r]);
    readers[t.src].push_back(kv_ready[r]);
  };

  for (int s = 0; s < W && R.rc == NNOP_OK; ++s) {
    std::vector<cudaEvent_t> ready_now = kv_ready;   // events of the block computed on in this step
    if (s + 1 < W)
      for (int r = 0; r < W; ++r) pull(r, s + 1);
    for (int r = 0; r < W && R.rc == NNOP_OK; ++r) {
      const StepShape t = step_shape(d, r, s);
      R.on(r);
      const void* kk = k[r];
      const void* vv = v[r];
      if (s > 0) {
        R.wait(R.comp[r], ready_now[r]);
        kk = ws(r, L.kv[s & 1]);
        vv = ws(r, L.kv[s & 1]) + kvb;
      }
      const void* qq = t.q_second ? static_cast<const void*>(ws(r, L.q2)) : q[r];
      if (int rc = nnop_flash_attn_fwd(ws(r, L.o_p), reinterpret_cast<float*>(ws(r, L.lse_p)), qq, kk, vv, nullptr,
                                       nullptr, dtype, E, t.QL, t.KL, QH, KH, B, t.causal, scale, R.comp[r]))
        return rc;
      RowMap m;
      m.rows = t.QL; m.acc_rows = d.Ll; m.acc_off = t.q_second ? d.c : 0;
      if (int rc = by_dtype(dtype, [&](auto tag) -> int {
            using T = decltype(tag);
            m.vpr = E / Vec<T>::N;
            m.nvec = static_cast<int64_t>(q_slabs) * t.QL * m.vpr;
            ring_merge_kernel<T><<<nblk(m.nvec), 256, 0, R.comp[r]>>>(
                reinterpret_cast<float*>(ws(r, L.o_acc)), lse[r], reinterpret_cast<const T*>(ws(r, L.o_p)),
                reinterpret_cast<const float*>(ws(r, L.lse_p)), m, s == 0);
            NNOP_LAUNCH_CHECK();
            return NNOP_OK;
          }))
        return rc;
      comp_done[r][s] = R.record(r, R.comp[r]);
    }
  }
  for (int r = 0; r < W && R.rc == NNOP_OK; ++r) {
    R.on(r);
    for (cudaEvent_t e : readers[r]) R.wait(R.comp[r], e);   // k[r] / v[r] stay live until every peer has its copy
    if (int rc = by_dtype(dtype, [&](auto tag) -> int {
          using T = decltype(tag);
          const int64_t nvec = static_cast<int64_t>(d.q_rows()) * (E / Vec<T>::N);
          ring_store_kernel<T><<<nblk(nvec), 256, 0, R.comp[r]>>>(static_cast<T*>(o[r]),
                                                                 reinterpret_cast<const float*>(ws(r, L.o_acc)), nvec);
          NNOP_LAUNCH_CHECK();
          return NNOP_OK;
        }))
      return rc;
  }
  return R.rc;
}

extern "C" int nnop_ring_attn_bwd(void* const* dq, void* const* dk, void* const* dv, const void* const* dO,
                                  const void* const* o, const float* const* lse, const void* const* q,
                                  const void* const* k, const void* const* v, const int* devices, int ndev,
                                  int dtype, int E, int Ll, int QH, int KH, int B, int causal, float scale,
                                  void* const* workspace, size_t workspace_bytes, void* const* streams) {
  clear_error();
  Dims d;
  if (int rc = check_dims(d, dtype, E, Ll, QH, KH, B, ndev, causal)) return rc;
  if (!devices) return fail(NNOP_ERR_ARG, "ring attention: NULL device list");
  if (int rc = check_ptrs(dq, ndev, "dq")) return rc;
  if (int rc = check_ptrs(dk, ndev, "dk")) return rc;
  if (int rc = check_ptrs(dv, ndev, "dv")) return rc;
  if (int rc = check_ptrs(dO, ndev, "dO")) return rc;
  if (int rc = check_ptrs(o, ndev, "o")) return rc;
  if (int rc = check_ptrs(lse, ndev, "lse")) return rc;
  if (int rc = check_ptrs(q, ndev, "q")) return rc;
  if (int rc = check_ptrs(k, ndev, "k")) return rc;
  if (int rc = check_ptrs(v, ndev, "v")) return rc;
  const Layout L = make_layout(d, true);
  if (workspace_bytes < L.total)
    return fail(NNOP_ERR_WORKSPACE, "ring attention backward needs a %zu-byte workspace per device, got %zu", L.total,
                workspace_bytes);
  if (!workspace) return fail(NNOP_ERR_WORKSPACE, "ring attention: NULL workspace array");
  for (int r = 0; r < ndev; ++r)
    if (!workspace[r] || (reinterpret_cast<uintptr_t>(workspace[r]) & 255) != 0)
      return fail(NNOP_ERR_WORKSPACE, "workspace[%d] must be non-NULL and 256-byte aligned", r);

  Ring R(d, devices, streams);
  if (R.init()) return R.rc;
  const int W = d.W;
  auto ws = [&](int r, size_t off) { return static_cast<char*>(workspace[r]) + off; };
  const size_t kvb = d.kv_full();
  const size_t kv_slabs = static_cast<size_t>(B) * KH, q_slabs = static_cast<size_t>(B) * QH;
  const size_t row_bytes = static_cast<size_t>(E) * d.es;

  std::vector<cudaEvent_t> in_ready(W), kv_ready(W, nullptr);
  std::vector<std::vector<cudaEvent_t>> readers(W);   // pulls that read rank r's k / v (other ranks' copy streams)
  std::vector<std::vector<cudaEvent_t>> comp_done(W, std::vector<cudaEvent_t>(W, nullptr)),
      part_ready(W, std::vector<cudaEvent_t>(W, nullptr)), sent(W, std::vector<cudaEvent_t>(W, nullptr)),
      acc_done(W, std::vector<cudaEvent_t>(W, nullptr));
  for (int r = 0; r < W; ++r) in_ready[r] = R.record(r, R.comp[r]);
  if (d.causal)
    for (int r = 0; r < W; ++r) {   // second-chunk copies of q, dO, o, lse
      R.on(r);
      R.copy_rows(ws(r, L.q2), r, q[r], r, q_slabs, d.c, d.c, R.comp[r]);
      R.copy_rows(ws(r, L.dO2), r, dO[r], r, q_slabs, d.c, d.c, R.comp[r]);
      R.copy_rows(ws(r, L.o2), r, o[r], r, q_slabs, d.c, d.c, R.comp[r]);
      R.set(cudaMemcpy2DAsync(ws(r, L.lse2), d.c * sizeof(float), lse[r] + d.c, d.Ll * sizeof(float),
                              d.c * sizeof(float), q_slabs, cudaMemcpyDeviceToDevice, R.comp[r]));
    }

  auto pull = [&](int r, int s) {
    const StepShape t = step_shape(d, r, s);
    R.on(r);
    R.wait(R.copy[r], in_ready[r]);
    R.wait(R.copy[r], in_ready[t.src]);
    if (s >= 3) R.wait(R.copy[r], comp_done[r][s - 2]);
    char* dst = ws(r, L.kv[s & 1]);
    R.copy_rows(dst, r, k[t.src], t.src, kv_slabs, 0, t.KL, R.copy[r]);
    R.copy_rows(dst + kvb, r, v[t.src], t.src, kv_slabs, 0, t.KL, R.copy[r]);
    kv_ready[r] = R.record(r, R.copy[r]);
    readers[t.src].push_back(kv_ready[r]);
  };
  auto accum = [&](int r, float* acc, const void* part, size_t slabs, int rows, int acc_off, int init) -> int {
    return by_dtype(dtype, [&](auto tag) -> int {
      using T = decltype(tag);
      RowMap m;
      m.vpr = E / Vec<T>::N; m.rows = rows; m.acc_rows = d.Ll; m.acc_off = acc_off;
      m.nvec = static_cast<int64_t>(slabs) * rows * m.vpr;
      ring_accum_kernel<T><<<nblk(m.nvec), 256, 0, R.comp[r]>>>(acc, static_cast<const T*>(part), m, init);
      NNOP_LAUNCH_CHECK();
      return NNOP_OK;
    });
  };

  for (int s = 0; s < W && R.rc == NNOP_OK; ++s) {
    std::vector<cudaEvent_t> ready_now = kv_ready;
    if (s + 1 < W)
      for (int r = 0; r < W; ++r) pull(r, s + 1);
    // ---- this step's pair on every rank: dq stays, [dk_p | dv_p] goes into part[s & 1]
    for (int r = 0; r < W && R.rc == NNOP_OK; ++r) {
      const StepShape t = step_shape(d, r, s);
      R.on(r);
      const void* kk = k[r];
      const void* vv = v[r];
      if (s > 0) {
        R.wait(R.comp[r], ready_now[r]);
        kk = ws(r, L.kv[s & 1]);
        vv = ws(r, L.kv[s & 1]) + kvb;
      }
      if (s >= 3) R.wait(R.comp[r], sent[r][s - 2]);   // part[s & 1] has left for its owner
      const void* qq = t.q_second ? static_cast<const void*>(ws(r, L.q2)) : q[r];
      const void* dd = t.q_second ? static_cast<const void*>(ws(r, L.dO2)) : dO[r];
      const void* oo = t.q_second ? static_cast<const void*>(ws(r, L.o2)) : o[r];
      const float* ll = t.q_second ? reinterpret_cast<const float*>(ws(r, L.lse2)) : lse[r];
      char* part = ws(r, L.part[s & 1]);
      const size_t part_half = kv_slabs * t.KL * row_bytes;   // dk_p, then dv_p right behind it
      if (int rc = nnop_flash_attn_bwd(ws(r, L.dq_p), part, part + part_half, nullptr, dd, oo, ll, qq, kk, vv, nullptr,
                                       nullptr, dtype, E, t.QL, t.KL, QH, KH, B, t.causal, scale, ws(r, L.attn_ws),
                                       L.attn_ws_bytes, R.comp[r]))
        return rc;
      if (int rc = accum(r, reinterpret_cast<float*>(ws(r, L.dq_acc)), ws(r, L.dq_p), q_slabs, t.QL,
                         t.q_second ? d.c : 0, s == 0))
        return rc;
      if (s == 0) {   // own block: straight into the accumulator (covers every row: initialises it)
        if (int rc = accum(r, reinterpret_cast<float*>(ws(r, L.dkv_acc)), part, 2 * kv_slabs, d.Ll, 0, 1)) return rc;
      } else {
        part_ready[r][s] = R.record(r, R.comp[r]);
      }
      comp_done[r][s] = R.record(r, R.comp[r]);
    }
    if (s == 0) continue;
    // ---- push the partials to the owners of the blocks
    for (int r = 0; r < W; ++r) {
      const StepShape t = step_shape(d, r, s);
      R.on(r);
      R.wait(R.copy[r], part_ready[r][s]);
      R.wait(R.copy[r], s >= 3 ? acc_done[t.src][s - 2] : in_ready[t.src]);   // owner's landing area is free
      const size_t bytes = 2 * kv_slabs * t.KL * row_bytes;
      R.set(cudaMemcpyPeerAsync(ws(t.src, L.recv[s & 1]), devices[t.src], ws(r, L.part[s & 1]), devices[r], bytes,
                                R.copy[r]));
      sent[r][s] = R.record(r, R.copy[r]);
    }
    // ---- owners add what arrived: the sender of step s is rank (owner + s) mod W
    for (int own = 0; own < W && R.rc == NNOP_OK; ++own) {
      const int sender = (own + s) % W;
      const StepShape t = step_shape(d, sender, s);   // t.src == own
      R.on(own);
      R.wait(R.comp[own], sent[sender][s]);
      // dk_p and dv_p arrive as (2 * slabs, KL, E); they address rows [0, KL) of the accumulator's slabs
      if (int rc = accum(own, reinterpret_cast<float*>(ws(own, L.dkv_acc)), ws(own, L.recv[s & 1]), 2 * kv_slabs, t.KL,
                         0, 0))
        return rc;
      acc_done[own][s] = R.record(own, R.comp[own]);
    }
  }
  for (int r = 0; r < W && R.rc == NNOP_OK; ++r) {
    R.on(r);
    for (int s = W - 2; s < W; ++s)   // the last pushes still read this rank's part buffers
      if (s >= 1) R.wait(R.comp[r], sent[r][s]);
    for (cudaEvent_t e : readers[r]) R.wait(R.comp[r], e);   // k[r] / v[r] stay live until every peer has its copy
    if (int rc = by_dtype(dtype, [&](auto tag) -> int {
          using T = decltype(tag);
          const int vpr = E / Vec<T>::N;
          const int64_t nq = static_cast<int64_t>(d.q_rows()) * vpr, nk = static_cast<int64_t>(d.kv_rows()) * vpr;
          const float* acc = reinterpret_cast<const float*>(ws(r, L.dkv_acc));
          ring_store_kernel<T><<<nblk(nq), 256, 0, R.comp[r]>>>(static_cast<T*>(dq[r]),
                                                               reinterpret_cast<const float*>(ws(r, L.dq_acc)), nq);
          ring_store_kernel<T><<<nblk(nk), 256, 0, R.comp[r]>>>(static_cast<T*>(dk[r]), acc, nk);
          ring_store_kernel<T><<<nblk(nk), 256, 0, R.comp[r]>>>(static_cast<T*>(dv[r]),
                                                               acc + d.kv_rows() * static_cast<size_t>(E), nk);
          NNOP_LAUNCH_CHECK();
          return NNOP_OK;
        }))
      return rc;
  }
  return R.rc;
}
