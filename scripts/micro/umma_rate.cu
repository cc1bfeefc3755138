// Microbenchmark (development aid, not part of the library): sustained issue rate of the
// tcgen05.mma operand forms the attention kernels use.  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nnop.jl_b200/csrc -o umma_rate scripts/micro/umma_rate.cu
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace nnop;

__global__ void __launch_bounds__(128, 1) k_rate(int form, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  // fill smem with small finite bf16 values
  for (int i = threadIdx.x; i < 4 * 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tslot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  if (threadIdx.x == 0) {
    const uint32_t a = smem_u32(smem), b = a + 32768, a2 = a + 65536, b2 = a + 98304;
    constexpr uint32_t id_kk128 = make_idesc_f16(128, 128, true, false, false);
    constexpr uint32_t id_kk64 = make_idesc_f16(128, 64, true, false, false);
    constexpr uint32_t id_kk256 = make_idesc_f16(128, 256, true, false, false);
    constexpr uint32_t id_tv = make_idesc_f16(128, 128, true, false, true);
    constexpr uint32_t id_mm = make_idesc_f16(128, 128, true, true, true);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (form == 0) {  // QK: SS, both K-major, N=128
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_ss(tb, make_smem_desc_sw128(a + off, 16, 1024), make_smem_desc_sw128(b + off, 16, 1024), id_kk128, 1);
        }
      } else if (form == 1) {  // PV: A in TMEM, B MN-major, N=128
        for (int j = 0; j < 8; ++j)
          umma_ts(tb + 256, tb + j * 8, make_smem_desc_sw128(b + j * 2048, 16384, 1024), id_tv, 1);
      } else if (form == 2) {  // SS N=64
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_ss(tb, make_smem_desc_sw128(a + off, 16, 1024), make_smem_desc_sw128(b + off, 16, 1024), id_kk64, 1);
        }
      } else if (form == 3) {  // SS N=256 (B spans two tiles)
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_ss(tb, make_smem_desc_sw128(a + off, 16, 1024), make_smem_desc_sw128(b + off, 16, 1024), id_kk256, 1);
        }
      } else if (form == 4) {  // QK then PV alternating (forward steady state)
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_ss(tb, make_smem_desc_sw128(a + off, 16, 1024), make_smem_desc_sw128(b + off, 16, 1024), id_kk128, 1);
        }
        for (int j = 0; j < 8; ++j)
          umma_ts(tb + 256, tb + 128 + j * 8, make_smem_desc_sw128(b2 + j * 2048, 16384, 1024), id_tv, 1);
      } else if (form == 5) {  // SS both MN-major (dQ form), N=128
        for (int ks = 0; ks < 8; ++ks)
          umma_ss(tb, make_smem_desc_sw128(a + ks * 2048, 16384, 1024), make_smem_desc_sw128(b + ks * 2048, 16384, 1024), id_mm, 1);
      } else if (form == 6) {  // SS A K-major, B MN-major (dK form)
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_ss(tb, make_smem_desc_sw128(a + off, 16, 1024), make_smem_desc_sw128(b + ks * 2048, 16384, 1024), id_tv, 1);
        }
      } else if (form == 7) {  // TS with B K-major N=128
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_ts(tb + 256, tb + ks * 8, make_smem_desc_sw128(b + off, 16, 1024), id_kk128, 1);
        }
      }
      (void)a2;
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

int main(int argc, char** argv) {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768);
  const char* names[] = {"SS K/K N=128 (QK)", "TS B=MN N=128 (PV)", "SS K/K N=64", "SS K/K N=256", "QK+PV alternating",
                         "SS MN/MN (dQ)", "SS K/MN (dK)", "TS B=K N=128"};
  const int iters = 2000;
  for (int grid : {1, 148}) {
    for (int form = 0; form < 8; ++form) {
      k_rate<<<grid, 128, 4 * 32768>>>(form, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("form %d: %s\n", form, cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const int n_mma = iters * (form == 4 ? 16 : 8);
      printf("grid %3d  %-22s %8.1f clk per UMMA (K=16)\n", grid, names[form], double(c) / n_mma);
    }
  }
  return 0;
}
