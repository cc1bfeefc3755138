// Microbenchmark (development aid): row-max reduction variants over 128 registers, one warp per SMSP.
#include "common.cuh"
#include <cstdio>
using namespace nnop;
__device__ __forceinline__ float vmax3(float a, float b, float c) { float d; asm volatile("max.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float vmax2(float a, float b) { float d; asm volatile("max.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ int vimax(int a, int b) { int d; asm volatile("max.s32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint64_t vadd2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE>
__global__ void __launch_bounds__(256, 1) k_max(const float* in, float* out, long long* clk, int reps, int nw) {
  if ((threadIdx.x >> 5) >= nw) return;
  float s[128];
#pragma unroll
  for (int j = 0; j < 128; ++j) s[j] = in[(threadIdx.x * 128 + j) % 4096];
  float tot = 0.f;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    float m8[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) m8[u] = -INFINITY;
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 64; ++j) m8[j & 7] = vmax3(m8[j & 7], s[2 * j], s[2 * j + 1]);
    } else if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 128; ++j) m8[j & 7] = vmax2(m8[j & 7], s[j]);
    } else if (MODE == 2) {  // integer max on the raw bits is wrong for negatives; timing only
#pragma unroll
      for (int j = 0; j < 128; ++j) m8[j & 7] = __int_as_float(vimax(__float_as_int(m8[j & 7]), __float_as_int(s[j])));
    } else if (MODE == 3) {  // packed: compare pairs with FFMA2-style? use add2 as a stand-in for cost of a packed op
      uint64_t a = pack_f2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 64; ++j) a = vadd2(a, pack_f2(s[2 * j], s[2 * j + 1]));
      float x, y; unpack_f2(a, x, y); m8[0] = x + y;
    }
    float m = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
    tot += m;
  }
  long long t1 = clock64();
  out[threadIdx.x] = tot;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  float* in; float* out; long long* clk;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 1024 * 4); cudaMalloc(&clk, 8);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = -0.001f * (i % 1000);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const char* names[] = {"64 x FMNMX3 (8 chains)", "128 x FMNMX (8 chains)", "128 x IMNMX (8 chains)", "64 x FADD2 (1 chain)"};
  for (int nw : {4, 8})
    for (int mode = 0; mode < 4; ++mode) {
      switch (mode) {
        case 0: k_max<0><<<1, 256>>>(in, out, clk, 200, nw); break;
        case 1: k_max<1><<<1, 256>>>(in, out, clk, 200, nw); break;
        case 2: k_max<2><<<1, 256>>>(in, out, clk, 200, nw); break;
        case 3: k_max<3><<<1, 256>>>(in, out, clk, 200, nw); break;
      }
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      printf("%d warps/SMSP  %-26s %7.1f clk per 128-element row\n", nw / 4, names[mode], double(c) / 200);
    }
  return 0;
}
