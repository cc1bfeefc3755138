// Microbenchmark (development aid): cta_group::2 tcgen05.mma (M=256 over a CTA pair, N=128, K=16).
// Per CTA an SS MMA should fetch A (own 128 rows, 4 KB) + half of B (64 rows, 2 KB) = 6 KB from shared
// memory instead of 8 KB: measures clk per MMA and the LSU bandwidth left over in each CTA.
#include "common.cuh"
#include <cstdio>
#include <cooperative_groups.h>
using namespace nnop;
namespace cg = cooperative_groups;

__device__ __forceinline__ void umma2_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void commit2(uint64_t* bar) {   // arrives on the same-offset barrier of both CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(288, 1)
k2(int form, int iters, int nload, long long* out, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t rank = cluster.block_rank();
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4 * 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
  fence_proxy_async_smem();
  cluster.sync();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster.sync();
  tc_fence_after();
  const uint32_t tb = tslot;
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    if (rank == 0) {
      const uint32_t a = smem_u32(smem), b = a + 32768;
      constexpr uint32_t id_ss = make_idesc_f16(256, 128, true, false, false);
      constexpr uint32_t id_ts = make_idesc_f16(256, 128, true, false, true);
      for (int it = 0; it < iters; ++it) {
        if (form == 0) {
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
            umma2_ss(tb, make_smem_desc_sw128(a + off, 16, 1024), make_smem_desc_sw128(b + off, 16, 1024), id_ss, 1);
          }
        } else {
          for (int j = 0; j < 8; ++j)
            umma2_ts(tb + 256, tb + j * 8, make_smem_desc_sw128(b + j * 2048, 16384, 1024), id_ts, 1);
        }
      }
      commit2(&bar);
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    stop = 1;
    if (blockIdx.x < 2) out[blockIdx.x * 32] = t1 - t0;
  } else if (warp >= 1 && warp <= nload) {
    const uint32_t base = smem_u32(smem) + 65536 + (threadIdx.x & 31) * 16;
    uint32_t acc = 0; long long n = 0;
    long long t0 = clock64();
    while (!stop) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        uint32_t x, y, z, w;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(base + ((warp * 16 + u) & 127) * 512));
        acc ^= x ^ y ^ z ^ w;
      }
      n += 16;
    }
    long long t1 = clock64();
    sink[threadIdx.x] = acc;
    if (blockIdx.x < 2 && (threadIdx.x & 31) == 0) { out[blockIdx.x * 32 + 2 * warp] = n; out[blockIdx.x * 32 + 2 * warp + 1] = t1 - t0; }
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
  }
}

int main() {
  long long* d; uint32_t* sink;
  cudaMalloc(&d, 8 * 64); cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768);
  const int iters = 2000;
  for (int form = 0; form < 2; ++form)
    for (int nload : {0, 4, 8}) {
      cudaMemset(d, 0, 8 * 64);
      k2<<<148, 288, 4 * 32768>>>(form, iters, nload, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("form %d: %s\n", form, cudaGetErrorString(e)); return 1; }
      long long h[64]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      for (int c = 0; c < 2; ++c) {
        double lds_bytes = 0, lds_clk = 1;
        for (int w = 1; w <= nload; ++w) { lds_bytes += h[c * 32 + 2 * w] * 512.0; lds_clk = h[c * 32 + 2 * w + 1]; }
        printf("%s M=256 N=128  load warps %d  CTA %d: %6.1f clk per MMA; LSU %.1f B/clk\n", form == 0 ? "2-CTA SS" : "2-CTA TS", nload, c,
               double(h[c * 32]) / (iters * 8), lds_bytes / lds_clk);
      }
    }
  return 0;
}
