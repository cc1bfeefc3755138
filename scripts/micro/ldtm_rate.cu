// Microbenchmark (development aid): tcgen05.ld / tcgen05.st throughput and latency per warp.
#include "common.cuh"
#include <cstdio>
using namespace nnop;
__device__ __forceinline__ uint32_t vxor(uint32_t a, uint32_t b) { uint32_t d; asm volatile("xor.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256, 1) k_ld(uint32_t* out, long long* clk, int reps, int nw) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nw) {
    uint32_t z[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) z[j] = threadIdx.x + j;
    for (int c = 0; c < 16; ++c) tmem_st_x32(tb + c * 32, z);
    tmem_st_wait();
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (MODE == 0) {  // 4 x (x32) then wait, consume one register of each
        uint32_t a[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_x32(tb + c * 32, a[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) acc = vxor(acc, a[c][31]);
      } else if (MODE == 1) {  // 1 x (x32) then wait: latency of one load
        uint32_t a[32];
        tmem_ld_x32(tb, a);
        tmem_ld_wait();
        acc = vxor(acc, a[31]);
      } else if (MODE == 2) {  // 4 x (x32) then consume ALL registers
        uint32_t a[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_x32(tb + c * 32, a[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 32; ++j) acc = vxor(acc, a[c][j]);
      } else if (MODE == 3) {  // store 4 x (x16) then wait (the P write)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t b[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) b[j] = z[j] + r;
          tmem_st_x16(tb + c * 16, b);
        }
        tmem_st_wait();
      } else if (MODE == 4) {  // 8 x (x16) loads
        uint32_t a[8][16];
#pragma unroll
        for (int c = 0; c < 8; ++c) tmem_ld_x16(tb + c * 16, a[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c) acc = vxor(acc, a[c][15]);
      }
    }
    t1 = clock64();
  }
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tslot); }
}
int main() {
  uint32_t* out; long long* clk;
  cudaMalloc(&out, 1024 * 4); cudaMalloc(&clk, 8);
  const char* names[] = {"ld 4 x32, wait, touch 4", "ld 1 x32, wait", "ld 4 x32, wait, touch 128", "st 4 x16, wait", "ld 8 x16, wait"};
  for (int nw : {1, 4, 8})
    for (int mode = 0; mode < 5; ++mode) {
      switch (mode) {
        case 0: k_ld<0><<<1, 256>>>(out, clk, 500, nw); break;
        case 1: k_ld<1><<<1, 256>>>(out, clk, 500, nw); break;
        case 2: k_ld<2><<<1, 256>>>(out, clk, 500, nw); break;
        case 3: k_ld<3><<<1, 256>>>(out, clk, 500, nw); break;
        case 4: k_ld<4><<<1, 256>>>(out, clk, 500, nw); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      printf("%d warps  %-28s %7.1f clk per iteration\n", nw, names[mode], double(c) / 500);
    }
  return 0;
}
