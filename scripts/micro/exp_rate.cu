// Microbenchmark (development aid): per-element cost of the softmax inner sequence on one warp per SMSP.
#include "common.cuh"
#include <cstdio>
using namespace nnop;

template <int MODE>
__global__ void __launch_bounds__(256, 1) k_exp(const float* in, float* out, long long* clk, int reps, int nwarps_active) {
  const int warp = threadIdx.x >> 5;
  if (warp >= nwarps_active) return;
  float s[128];
#pragma unroll
  for (int j = 0; j < 128; ++j) s[j] = in[(threadIdx.x * 128 + j) % 4096];
  float sl2 = in[0], m = in[1];
  uint32_t acc = 0; float sum = 0.f;
  uint64_t sum2 = pack_f2(0.f, 0.f);
  const uint64_t sl2x2 = pack_f2(sl2, sl2), negm2 = pack_f2(-m, -m);
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      if (MODE == 0) {  // MUFU only
        float p0 = fast_exp2(s[2 * j]), p1 = fast_exp2(s[2 * j + 1]);
        s[2 * j] = p0 - 1.0f; s[2 * j + 1] = p1 - 1.0f;
      } else if (MODE == 1) {  // MUFU + F2FP
        float p0 = fast_exp2(s[2 * j]), p1 = fast_exp2(s[2 * j + 1]);
        acc ^= pack2<__nv_bfloat16>(p0, p1);
        s[2 * j] = p0 - 1.0f; s[2 * j + 1] = p1 - 1.0f;  // feedback: keeps every rep live
      } else if (MODE == 2) {  // full: FFMA2 + 2 MUFU + FADD2 + F2FP
        float x0, x1;
        unpack_f2(ffma2(pack_f2(s[2 * j], s[2 * j + 1]), sl2x2, negm2), x0, x1);
        float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
        sum2 = fadd2(sum2, pack_f2(p0, p1));
        acc ^= pack2<__nv_bfloat16>(p0, p1);
        s[2 * j] = p0 - 1.0f; s[2 * j + 1] = p1 - 1.0f;  // feedback: keeps every rep live
      } else if (MODE == 3) {  // full with every 4th pair on the FMA pipe
        float x0, x1, p0, p1;
        unpack_f2(ffma2(pack_f2(s[2 * j], s[2 * j + 1]), sl2x2, negm2), x0, x1);
        if ((j & 3) == 3) exp2_poly2(x0, x1, p0, p1); else { p0 = fast_exp2(x0); p1 = fast_exp2(x1); }
        sum2 = fadd2(sum2, pack_f2(p0, p1));
        acc ^= pack2<__nv_bfloat16>(p0, p1);
        s[2 * j] = p0 - 1.0f; s[2 * j + 1] = p1 - 1.0f;  // feedback: keeps every rep live
      } else if (MODE == 4) {  // every 2nd pair
        float x0, x1, p0, p1;
        unpack_f2(ffma2(pack_f2(s[2 * j], s[2 * j + 1]), sl2x2, negm2), x0, x1);
        if ((j & 1) == 1) exp2_poly2(x0, x1, p0, p1); else { p0 = fast_exp2(x0); p1 = fast_exp2(x1); }
        sum2 = fadd2(sum2, pack_f2(p0, p1));
        acc ^= pack2<__nv_bfloat16>(p0, p1);
        s[2 * j] = p0 - 1.0f; s[2 * j + 1] = p1 - 1.0f;  // feedback: keeps every rep live
      } else if (MODE == 5) {  // F2FP only
        acc ^= pack2<__nv_bfloat16>(s[2 * j], s[2 * j + 1]);
        s[2 * j] += 1.0f;
      } else if (MODE == 6) {  // all poly
        float x0, x1, p0, p1;
        unpack_f2(ffma2(pack_f2(s[2 * j], s[2 * j + 1]), sl2x2, negm2), x0, x1);
        exp2_poly2(x0, x1, p0, p1);
        sum2 = fadd2(sum2, pack_f2(p0, p1));
        acc ^= pack2<__nv_bfloat16>(p0, p1);
        s[2 * j] = p0 - 1.0f; s[2 * j + 1] = p1 - 1.0f;  // feedback: keeps every rep live
      }
    }
  }
  long long t1 = clock64();
  float a, b; unpack_f2(sum2, a, b); sum += a + b;
#pragma unroll
  for (int j = 0; j < 128; ++j) sum += s[j];
  out[threadIdx.x] = sum + __uint_as_float(acc);
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}

int main() {
  float* in; float* out; long long* clk;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 1024 * 4); cudaMalloc(&clk, 8);
  float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = -0.001f * (i % 1000); h[0] = 0.1f; h[1] = 0.5f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const char* names[] = {"MUFU only", "MUFU+F2FP", "FFMA2+MUFU+FADD2+F2FP", "full, 1/4 poly", "full, 1/2 poly", "F2FP only", "all poly"};
  const int reps = 200;
  for (int nw : {4, 8}) {
    for (int mode = 0; mode < 7; ++mode) {
      switch (mode) {
        case 0: k_exp<0><<<1, 256>>>(in, out, clk, reps, nw); break;
        case 1: k_exp<1><<<1, 256>>>(in, out, clk, reps, nw); break;
        case 2: k_exp<2><<<1, 256>>>(in, out, clk, reps, nw); break;
        case 3: k_exp<3><<<1, 256>>>(in, out, clk, reps, nw); break;
        case 4: k_exp<4><<<1, 256>>>(in, out, clk, reps, nw); break;
        case 5: k_exp<5><<<1, 256>>>(in, out, clk, reps, nw); break;
        case 6: k_exp<6><<<1, 256>>>(in, out, clk, reps, nw); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      printf("%d warps/SMSP  %-24s %6.2f clk per element per warp (%.0f clk per 128-element row)\n", nw / 4, names[mode],
             double(c) / (reps * 128.0), double(c) / reps);
    }
  }
  return 0;
}
