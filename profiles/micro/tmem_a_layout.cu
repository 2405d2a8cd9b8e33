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

constexpr int M = 128, N = 64, K = 16;

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

__global__ void __launch_bounds__(128) k_test(const __nv_bfloat16 *A, const __nv_bfloat16 *B,
                                              float *D) {
  __shared__ __align__(1024) uint8_t sB[N * K * 2];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // B: [N rows][16 channels] K-major, 32 B rows, 32-byte swizzle (Swz<16> of conv_tc.cu)
  for (uint32_t e = tid; e < N * 2; e += 128) {
    const uint32_t row = e >> 1, chunk = e & 1;
    const uint32_t x = (row >> 2) & 1u;
    *reinterpret_cast<uint4 *>(sB + row * 32 + ((chunk ^ x) << 4)) =
        *reinterpret_cast<const uint4 *>(B + row * K + chunk * 8);
  }
  fence_proxy_async();
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 128); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t tmem_d = tmem_base, tmem_a = tmem_base + 64;      // D: 64 columns, A: 8 columns
  // A: thread = row (TMEM lane), 8 registers = 16 bf16
  uint32_t r[8];
  const uint32_t *arow = reinterpret_cast<const uint32_t *>(A + (size_t)tid * K);
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = arow[i];
  tmem_st8(tmem_a + ((warp * 32u) << 16), r);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(1 /*bf16*/, M, N, 0, 0);
    const uint64_t db = umma_desc(smem_u32(sB), 16, 8 * 32, 6 /*SWIZZLE_32B*/);
    umma_f16_ts(tmem_d, tmem_a, db, idesc, 0);
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
  cudaMemset(dD, 0, sizeof(hD));
  k_test<<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
  double worst = 0; int bad = 0;
  for (int i = 0; i < M * N; ++i) {
    double d = fabs((double)hD[i] - ref[i]);
    if (d > worst) worst = d;
    if (d > 1e-3) ++bad;
  }
  printf("A-in-TMEM (lane = row, column = bf16 pair): max abs diff %.3g, %d of %d wrong -> %s\n",
         worst, bad, M * N, bad == 0 ? "TMEM_A_LAYOUT_OK" : "TMEM_A_LAYOUT_MISMATCH");
  if (bad) {
    printf("row 0, first 8 outputs: got");
    for (int j = 0; j < 8; ++j) printf(" %.3f", hD[j]);
    printf(" | want");
    for (int j = 0; j < 8; ++j) printf(" %.3f", ref[j]);
    printf("\n");
  }
  return 0;
}
