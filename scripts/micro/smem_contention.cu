// Microbenchmark (development aid): does tcgen05.mma operand fetch share the 128 B/clk shared-memory
// port with LSU traffic?  One thread streams SS / TS MMAs while `nload` warps stream ld.shared.v4.
#include "common.cuh"
#include <cstdio>
using namespace nnop;

__global__ void __launch_bounds__(288, 1) k_cont(int form, int iters, int nload, int lds_per_mma_iter, long long* out, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4 * 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
  if (warp == 0) tmem_alloc<512>(&tslot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = a + 32768;
    constexpr uint32_t id_kk128 = make_idesc_f16(128, 128, true, false, false);
    constexpr uint32_t id_tv = make_idesc_f16(128, 128, true, false, true);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (form == 0) {
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_ss(tb, make_smem_desc_sw128(a + off, 16, 1024), make_smem_desc_sw128(b + off, 16, 1024), id_kk128, 1);
        }
      } else if (form == 1) {
        for (int j = 0; j < 8; ++j)
          umma_ts(tb + 256, tb + j * 8, make_smem_desc_sw128(b + j * 2048, 16384, 1024), id_tv, 1);
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    stop = 1;
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (warp >= 1 && warp <= nload) {
    // stream conflict-free 16-byte loads over a 64 KB window (upper half of the buffer)
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
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[2 * warp] = n; out[2 * warp + 1] = t1 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

int main() {
  long long* d; uint32_t* sink;
  cudaMalloc(&d, 8 * 32); cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(k_cont, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768);
  const int iters = 2000;
  for (int form = 0; form < 2; ++form)
    for (int nload : {0, 1, 2, 4, 8}) {
      cudaMemset(d, 0, 8 * 32);
      k_cont<<<148, 288, 4 * 32768>>>(form, iters, nload, 0, d, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
      long long h[32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double lds_bytes = 0, lds_clk = 1;
      for (int w = 1; w <= nload; ++w) { lds_bytes += h[2 * w] * 512.0; lds_clk = h[2 * w + 1]; }
      printf("%s  load warps %d: %6.1f clk per UMMA; LSU %.1f B/clk; tensor-operand %.1f B/clk\n", form == 0 ? "SS (8 KB/UMMA)" : "TS (4 KB/UMMA)", nload,
             double(h[0]) / (iters * 8), lds_bytes / lds_clk, (form == 0 ? 8192.0 : 4096.0) / (double(h[0]) / (iters * 8)));
    }
  return 0;
}
