// Microbenchmark (development aid): fp32 add-reduction of 128x128 tiles into global memory.
//   A: red.global.add.v4.f32, lane <-> row (what a tcgen05.ld 32x32b register tile gives directly)
//   B: red.global.add.v4.f32, a warp instruction covers 512 contiguous bytes of one row
//   C: st.shared + cp.reduce.async.bulk (TMA-style bulk reduce from smem, 16 KB pieces)
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace nnop;
__device__ __forceinline__ void red_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__global__ void __launch_bounds__(128, 1) k_red(int mode, float* buf, long long ntiles, int T, long long* clk) {
  __shared__ __align__(1024) float stage[2][128 * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float v[8];
  for (int j = 0; j < 8; ++j) v[j] = 1e-3f * (j + 1);
  long long t0 = clock64();
  int nred = 0;
  for (int t = 0; t < T; ++t) {
    const long long tile = ((long long)(blockIdx.x * T + t) * 7919) % ntiles;
    float* base = buf + tile * (128 * 128);
    if (mode == 0) {
#pragma unroll 8
      for (int c = 0; c < 32; ++c) red_v4(base + threadIdx.x * 128 + c * 4, v[c & 7], v[1], v[2], v[3]);
    } else if (mode == 1) {
#pragma unroll 8
      for (int r = 0; r < 32; ++r) red_v4(base + (warp * 32 + r) * 128 + lane * 4, v[r & 7], v[1], v[2], v[3]);
    } else if (mode == 3) {
      // tcgen05.ld 16x256b fragment layout: lane t holds rows (t/4) and (t/4)+8 of a 16-row slab,
      // columns 2*(t%4)+{0,1} of every 8-column group: a quad covers one 32-byte sector per red.v2
#pragma unroll 4
      for (int slab = 0; slab < 2; ++slab)      // warp owns 32 rows = 2 slabs of 16
#pragma unroll 4
        for (int g = 0; g < 16; ++g)            // 16 column groups of 8
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
            red_v2(base + (warp * 32 + slab * 16 + hh * 8 + (lane >> 2)) * 128 + g * 8 + (lane & 3) * 2, v[g & 7], v[1]);
    } else if (mode == 4) {
      // 16x128b-like: lane t holds row (t/2)... emulate: pairs of lanes cover 32 B with v4 each (16 rows x 32 B per instr)
#pragma unroll 4
      for (int slab = 0; slab < 2; ++slab)
#pragma unroll 4
        for (int g = 0; g < 16; ++g)
          red_v4(base + (warp * 32 + slab * 16 + (lane >> 1)) * 128 + g * 8 + (lane & 1) * 4, v[g & 7], v[1], v[2], v[3]);
    } else {
      for (int c = 0; c < 4; ++c) {  // 4 pieces of 128 rows x 32 floats = 16 KB, contiguous destination for simplicity
        float* st = stage[nred & 1];
        if (threadIdx.x == 0) bulk_wait_read<1>();
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 8; ++u)
          *reinterpret_cast<float4*>(st + threadIdx.x * 32 + ((u ^ (threadIdx.x & 7)) << 2)) = make_float4(v[u], v[1], v[2], v[3]);
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) { bulk_reduce_add_f32(base + c * 4096, st, 16384); bulk_commit(); }
        ++nred;
      }
    }
  }
  if (mode == 2 && threadIdx.x == 0) bulk_wait<0>();
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) clk[0] = t1 - t0;
}
int main(int argc, char** argv) {
  const long long ntiles = (argc > 1 ? atoll(argv[1]) : 16384);
  float* buf; long long* clk;
  cudaMalloc(&buf, ntiles * 65536); cudaMalloc(&clk, 8);
  cudaMemset(buf, 0, ntiles * 65536);
  const char* names[] = {"red.v4 lane<->row", "red.v4 row-contiguous", "smem + bulk reduce", "red.v2 quad=sector (16x256b)", "red.v4 pair=sector"};
  const int T = 1000;
  for (int mode = 0; mode < 5; ++mode) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k_red<<<148, 128>>>(mode, buf, ntiles, 50, clk);
    cudaEventRecord(a);
    k_red<<<148, 128>>>(mode, buf, ntiles, T, clk);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("%-24s %8.1f clk per 64 KB tile per SM, %7.1f GB/s chip-wide\n", names[mode], double(c) / T, 148.0 * T * 65536 / ms / 1e6);
  }
  return 0;
}
