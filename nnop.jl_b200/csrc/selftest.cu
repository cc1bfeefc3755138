// selftest.cu -- 128x128x128 GEMMs through each tcgen05 operand form the attention kernels
// use, so a descriptor / layout mistake shows up in isolation (tests/test_umma_selftest.py):
//   which 0  D = A  B^T   A, B via TMA (SWIZZLE_128B), both K-major from smem        (S = Q K^T)
//   which 1  D = A  B     A packed into TMEM by tcgen05.st, B via TMA, MN-major     (O = P V)
//   which 2  D = A  B^T   as 0 but A, B written to smem by threads with the manual
//                         128-byte-swizzle formula                                   (dS staging)
//   which 3  D = A^T B    A, B via TMA, both MN-major from smem                      (dQ = dS K)
//   which 4  D = A  B     A via TMA K-major, B via TMA MN-major                      (dK = dS^T Q)
// a, b are 128x128 row-major 16-bit (bf16) matrices; d_out is 128x128 fp32 row-major.
#include "common.cuh"
#include "internal.h"

namespace nnop {
namespace {

constexpr int kBox = 128 * 64 * 2;

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const __grid_constant__ CUtensorMap tm_a,
                     const __grid_constant__ CUtensorMap tm_b, const __nv_bfloat16* a,
                     const __nv_bfloat16* b, float* d_out, int which) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * kBox;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * kBox);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = threadIdx.x;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tD = tmem_base, tP = tmem_base + 128;

  if (which == 2) {
    // manual swizzled store of both operands
    for (int c = 0; c < 16; ++c) {
      const uint4 va = *reinterpret_cast<const uint4*>(a + row * 128 + c * 8);
      const uint4 vb = *reinterpret_cast<const uint4*>(b + row * 128 + c * 8);
      const int off = (c >> 3) * kBox + row * 128 + (((c & 7) ^ (row & 7)) << 4);
      *reinterpret_cast<uint4*>(sA + off) = va;
      *reinterpret_cast<uint4*>(sB + off) = vb;
    }
    fence_proxy_async_smem();
  } else if (threadIdx.x == 0) {
    const bool need_a = which != 1;
    mbar_arrive_expect_tx(&bars[0], (need_a ? 4 : 2) * kBox);
    for (int bx = 0; bx < 2; ++bx) {
      if (need_a) tma_load_3d(sA + bx * kBox, &tm_a, &bars[0], bx * 64, 0, 0);
      tma_load_3d(sB + bx * kBox, &tm_b, &bars[0], bx * 64, 0, 0);
    }
  }
  if (which == 1) {
    // A (row-major, K contiguous) -> TMEM: lane = row, 32-bit column j = elements 2j, 2j+1
    uint32_t pr[2][32];
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(a + row * 128);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      pr[0][j] = arow[j];
      pr[1][j] = arow[32 + j];
    }
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    tmem_st_x32(tP + lane_off, pr[0]);
    tmem_st_x32(tP + lane_off + 32, pr[1]);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    if (which != 2) mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    if (which == 0 || which == 2) {
      constexpr uint32_t idesc = make_idesc_f16(128, 128, true, false, false);
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t off = (ks >> 2) * kBox + (ks & 3) * 32;
        umma_ss(tD, make_smem_desc_sw128(a0 + off, 16, 1024),
                make_smem_desc_sw128(b0 + off, 16, 1024), idesc, ks > 0);
      }
    } else if (which == 1) {
      constexpr uint32_t idesc = make_idesc_f16(128, 128, true, false, true);
      for (int j = 0; j < 8; ++j)
        umma_ts(tD, tP + j * 8, make_smem_desc_sw128(b0 + j * 2048, kBox, 1024), idesc, j > 0);
    } else if (which == 3) {
      constexpr uint32_t idesc = make_idesc_f16(128, 128, true, true, true);
      for (int j = 0; j < 8; ++j)
        umma_ss(tD, make_smem_desc_sw128(a0 + j * 2048, kBox, 1024),
                make_smem_desc_sw128(b0 + j * 2048, kBox, 1024), idesc, j > 0);
    } else {
      constexpr uint32_t idesc = make_idesc_f16(128, 128, true, false, true);
      for (int ks = 0; ks < 8; ++ks) {
        const uint32_t off = (ks >> 2) * kBox + (ks & 3) * 32;
        umma_ss(tD, make_smem_desc_sw128(a0 + off, 16, 1024),
                make_smem_desc_sw128(b0 + ks * 2048, kBox, 1024), idesc, ks > 0);
      }
    }
    tc_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld_x32(tD + lane_off + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) d_out[row * 128 + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
  (void)lane;
}

}  // namespace
}  // namespace nnop

using namespace nnop;

extern "C" int nnop_selftest_umma(float* d_out, const void* a, const void* b, int which,
                                  void* stream) {
  clear_error();
  if (!d_out || !a || !b) return fail(NNOP_ERR_ARG, "NULL pointer");
  if (which < 0 || which > 4) return fail(NNOP_ERR_ARG, "which must be in [0, 4]");
  alignas(64) CUtensorMap ta, tb;
  if (int rc = make_tmap_3d(&ta, a, NNOP_BF16, 128, 128, 1, 64, 128)) return rc;
  if (int rc = make_tmap_3d(&tb, b, NNOP_BF16, 128, 128, 1, 64, 128)) return rc;
  const int smem = 4 * kBox + 64 + 1024;
  NNOP_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      ta, tb, static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), d_out,
      which);
  NNOP_LAUNCH_CHECK();
  return NNOP_OK;
}
