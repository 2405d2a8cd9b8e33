// Micro-test for the "operand A in tensor memory" design candidate (profiles/r1_notes.md):
// D[128 x 64] = A[128 x 16] * B[64 x 16]^T with A written to TMEM by tcgen05.st (one thread =
// one row = one TMEM lane, 8 columns of packed bf16 pairs) and B in shared memory (K-major,
// 32-byte swizzle, the layout the product kernels use for 16-channel chunks).  Prints whether
// the hypothesis "lane = row, column c = elements (2c, 2c+1)" reproduces the CPU product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -std=c++17 \
//        -o profiles/micro/tmem_a_layout profiles/micro/tmem_a_layout.cu && profiles/micro/tmem_a_layout
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../minkowskiengine_b200/csrc/ptx.cuh"
using namespace meb200::ptx;

constexpr int M = 128, N = 64, K = 32;   // two MMA K-steps of 16

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),
                 "r"(r[6]), "r"(r[7]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D(tmem) (+)= A(tmem) * B(smem descriptor)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// 16 lanes x 256 bits, two column groups: thread t holds lanes t/4 (regs 0,1,4,5) and t/4+8
// (regs 2,3,6,7); regs (0,1) = columns 2(t%4), 2(t%4)+1 of group 0, regs (4,5) the same of group 1
// (CuTe: SM100_TMEM_STORE_16dp256b2x, copy_traits_sm100.hpp)
__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]),
                 "r"(r[6]), "r"(r[7]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

template <int MODE>
__global__ void __launch_bounds__(128) k_test(const __nv_bfloat16 *A, const __nv_bfloat16 *B,
                                              float *D) {
  __shared__ __align__(1024) uint8_t sB[N * K * 2];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // B: [N rows][32 channels] K-major, 64 B rows, 64-byte swizzle (Swz<32> of conv_tc.cu)
  for (uint32_t e = tid; e < N * 4; e += 128) {
    const uint32_t row = e >> 2, chunk = e & 3;
    const uint32_t x = (row >> 1) & 3u;
    *reinterpret_cast<uint4 *>(sB + row * 64 + ((chunk ^ x) << 4)) =
        *reinterpret_cast<const uint4 *>(B + row * K + chunk * 8);
  }
  fence_proxy_async();
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tmem_d = tmem_base, tmem_a = tmem_base + 64;      // D: 64 columns, A: 16 columns
  if constexpr (MODE == 0) {
    // A: thread = row (TMEM lane), 16 registers = 32 bf16, two stores of 8 columns
    const uint32_t *arow = reinterpret_cast<const uint32_t *>(A + (size_t)tid * K);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = arow[h * 8 + i];
      tmem_st8(tmem_a + h * 8 + ((warp * 32u) << 16), r);
    }
  } else {
    // A: four threads per row; thread q = lane & 3 holds columns 2q, 2q+1 of each 8-column group
#pragma unroll
    for (int half = 0; half < 2; ++half) {          // lanes [16 half, 16 half + 16) of the quadrant
      uint32_t r[8];
      const uint32_t q = lane & 3u;
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {              // rows lane/4 and lane/4 + 8
        const uint32_t row = warp * 32 + half * 16 + hi * 8 + (lane >> 2);
        const uint32_t *arow = reinterpret_cast<const uint32_t *>(A + (size_t)row * K);
#pragma unroll
        for (int g = 0; g < 2; ++g) {               // column groups (one per MMA K-step)
          r[g * 4 + hi * 2 + 0] = arow[g * 8 + 2 * q];
          r[g * 4 + hi * 2 + 1] = arow[g * 8 + 2 * q + 1];
        }
      }
      tmem_st_16x256b_x2(tmem_a + (((warp * 32u + half * 16u)) << 16), r);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(1 /*bf16*/, M, N, 0, 0);
    for (uint32_t j = 0; j < 2; ++j) {
      const uint64_t db = umma_desc(smem_u32(sB) + j * 32, 16, 8 * 64, 4 /*SWIZZLE_64B*/);
      umma_f16_ts(tmem_d, tmem_a + j * 8, db, idesc, j);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (uint32_t c = 0; c < N; c += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_d + ((warp * 32u) << 16) + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
  (void)lane;
}

int main() {
  static __nv_bfloat16 hA[M * K], hB[N * K];
  static float hD[M * N], ref[M * N];
  srand(1);
  for (int i = 0; i < M * K; ++i) hA[i] = __float2bfloat16((rand() % 17 - 8) / 8.0f);
  for (int i = 0; i < N * K; ++i) hB[i] = __float2bfloat16((rand() % 13 - 6) / 4.0f);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += __bfloat162float(hA[i * K + k]) * __bfloat162float(hB[j * K + k]);
      ref[i * N + j] = s;
    }
  __nv_bfloat16 *dA, *dB;
  float *dD;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 2; ++mode) {
  cudaMemset(dD, 0, sizeof(hD));
  if (mode == 0) k_test<0><<<1, 128>>>(dA, dB, dD);
  else k_test<1><<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  double worst = 0; int bad = 0;
  for (int i = 0; i < M * N; ++i) {
    double d = fabs((double)hD[i] - ref[i]);
    if (d > worst) worst = d;
    if (d > 1e-3) ++bad;
  }
  printf("A-in-TMEM mode %d (%s; lane = row, column = bf16 pair): max abs diff %.3g, %d of %d wrong -> %s\n",
         mode, mode == 0 ? "tcgen05.st 32x32b, thread = row" : "tcgen05.st 16x256b.x2, 4 threads per row",
         worst, bad, M * N, bad == 0 ? "TMEM_A_LAYOUT_OK" : "TMEM_A_LAYOUT_MISMATCH");
  if (bad) {
    printf("row 0, first 8 outputs: got");
    for (int j = 0; j < 8; ++j) printf(" %.3f", hD[j]);
    printf(" | want");
    for (int j = 0; j < 8; ++j) printf(" %.3f", ref[j]);
    printf("\n");
  }
  }
  return 0;
}
