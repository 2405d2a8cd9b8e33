// Microbenchmark: how fast can one SM pull gathered rows into shared memory on B200?
//   mode 0: cp.async.cg 16 B (LDGSTS), 8 lanes per 128-B row
//   mode 1: ld.global.v4 -> st.shared.v4 (register staged, UNROLL loads in flight per thread)
//   mode 2: TMA tile::gather4 (one instruction = 4 rows), issued by all lanes of one warp
//   mode 3: cp.async.bulk (UBLKCP) one row per instruction, linear destination
// Rows are ROW_BYTES wide, picked at random from a table larger than L2 (or resident, see -r).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROW_BYTES>
__global__ void __launch_bounds__(1024, 1)
k_cpasync(const uint8_t *__restrict__ tab, const int32_t *__restrict__ idx, int rows_per_cta, int valid_pct) {
  extern __shared__ __align__(1024) uint8_t sm[];
  constexpr int CH = ROW_BYTES / 16;
  const int lane_row = threadIdx.x / CH, chunk = threadIdx.x % CH;
  const int rows_per_iter = blockDim.x / CH;
  const int32_t *my = idx + (size_t)blockIdx.x * rows_per_cta;
  const int ring_rows = (200 * 1024) / ROW_BYTES;
  int slot = 0;
  for (int r0 = 0; r0 < rows_per_cta; r0 += rows_per_iter) {
    int32_t i = my[r0 + lane_row];
    uint32_t nbytes = ((i & 127) * 100 < valid_pct * 128) ? 16u : 0u;
    const uint8_t *src = tab + (size_t)i * ROW_BYTES + chunk * 16;
    uint32_t dst = smem_u32(sm) + ((slot + lane_row) % ring_rows) * ROW_BYTES + chunk * 16;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 8;" ::: "memory");
    slot = (slot + rows_per_iter) % ring_rows;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int ROW_BYTES, int UNROLL>
__global__ void __launch_bounds__(1024, 1)
k_ldgsts(const uint8_t *__restrict__ tab, const int32_t *__restrict__ idx, int rows_per_cta, int valid_pct) {
  extern __shared__ __align__(1024) uint8_t sm[];
  constexpr int CH = ROW_BYTES / 16;
  const int lane_row = threadIdx.x / CH, chunk = threadIdx.x % CH;
  const int rows_per_iter = blockDim.x / CH;
  const int32_t *my = idx + (size_t)blockIdx.x * rows_per_cta;
  const int ring_rows = (200 * 1024) / ROW_BYTES;
  int slot = 0;
  for (int r0 = 0; r0 < rows_per_cta; r0 += rows_per_iter * UNROLL) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      int32_t i = my[r0 + u * rows_per_iter + lane_row];
      bool ok = ((i & 127) * 100 < valid_pct * 128);
      v[u] = ok ? *reinterpret_cast<const uint4 *>(tab + (size_t)i * ROW_BYTES + chunk * 16) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      *reinterpret_cast<uint4 *>(sm + ((slot + u * rows_per_iter + lane_row) % ring_rows) * ROW_BYTES + chunk * 16) = v[u];
    }
    slot = (slot + rows_per_iter * UNROLL) % ring_rows;
  }
}

template <int ROW_BYTES>
__global__ void __launch_bounds__(128, 1)
k_tma_gather4(const __grid_constant__ CUtensorMap map, const int32_t *__restrict__ idx, int rows_per_cta,
              int valid_pct, int n_rows_tab) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t *sm = (uint8_t *)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp != 0) return;
  const int32_t *my = idx + (size_t)blockIdx.x * rows_per_cta;
  // stage = 128 rows (32 lanes x gather4); 8 stages in flight
  const int n_stages = rows_per_cta / 128;
  for (int s = 0; s < n_stages; ++s) {
    const int b = s & 7;
    if (s >= 8) {
      uint32_t ok = 0, par = ((s >> 3) - 1) & 1;
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar[b])), "r"(par) : "memory");
    }
    int32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int32_t i = my[s * 128 + lane * 4 + j];
      r[j] = ((i & 127) * 100 < valid_pct * 128) ? i : n_rows_tab;
    }
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(128 * ROW_BYTES) : "memory");
    __syncwarp();
    uint32_t dst = smem_u32(sm) + b * 128 * ROW_BYTES + lane * 4 * ROW_BYTES;
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(&map), "r"(0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(smem_u32(&bar[b])) : "memory");
  }
  // drain
  for (int s = (n_stages > 8 ? n_stages - 8 : 0); s < n_stages; ++s) {
    const int b = s & 7; uint32_t ok = 0, par = (s >> 3) & 1;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar[b])), "r"(par) : "memory");
  }
}

template <int ROW_BYTES>
__global__ void __launch_bounds__(128, 1)
k_bulk_rows(const uint8_t *__restrict__ tab, const int32_t *__restrict__ idx, int rows_per_cta, int valid_pct) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t *sm = (uint8_t *)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp != 0) return;
  const int32_t *my = idx + (size_t)blockIdx.x * rows_per_cta;
  const int n_stages = rows_per_cta / 128;
  for (int s = 0; s < n_stages; ++s) {
    const int b = s & 7;
    if (s >= 8) {
      uint32_t ok = 0, par = ((s >> 3) - 1) & 1;
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar[b])), "r"(par) : "memory");
    }
    int nvalid = 0;
    int32_t r[4]; bool ok4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      r[j] = my[s * 128 + lane * 4 + j];
      ok4[j] = ((r[j] & 127) * 100 < valid_pct * 128);
      nvalid += ok4[j];
    }
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(nvalid * ROW_BYTES) : "memory");
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (!ok4[j]) continue;
      uint32_t dst = smem_u32(sm) + b * 128 * ROW_BYTES + (lane * 4 + j) * ROW_BYTES;
      asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(tab + (size_t)r[j] * ROW_BYTES), "r"(ROW_BYTES), "r"(smem_u32(&bar[b])) : "memory");
    }
  }
  for (int s = (n_stages > 8 ? n_stages - 8 : 0); s < n_stages; ++s) {
    const int b = s & 7; uint32_t ok = 0, par = (s >> 3) & 1;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(&bar[b])), "r"(par) : "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROW_BYTES>
void run(int n_rows_tab, int rows_per_cta, int valid_pct, const char *label) {
  const int n_cta = 148;
  uint8_t *tab; int32_t *idx;
  CK(cudaMalloc(&tab, (size_t)n_rows_tab * ROW_BYTES));
  CK(cudaMemset(tab, 1, (size_t)n_rows_tab * ROW_BYTES));
  std::vector<int32_t> h((size_t)n_cta * rows_per_cta);
  uint64_t s = 88172645463325252ull;
  for (auto &v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (int32_t)(s % n_rows_tab); }
  CK(cudaMalloc(&idx, h.size() * 4));
  CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const size_t smem = 201 * 1024 + 1024;
  auto report = [&](const char *mode, float ms) {
    double bytes = (double)n_cta * rows_per_cta * ROW_BYTES;
    double clk = 1.9e9;
    printf("%-26s row=%3dB valid=%3d%% %-12s %8.3f ms  %7.1f GB/s filled  %6.1f B/cyc/SM (valid %.1f)\n", label, ROW_BYTES, valid_pct, mode, ms,
           bytes / ms / 1e6, bytes / (ms * 1e-3 * clk) / n_cta, bytes * valid_pct / 100 / (ms * 1e-3 * clk) / n_cta);
  };
  float ms;
  for (int threads : {128, 256, 512, 1024}) {
    CK(cudaFuncSetAttribute(k_cpasync<ROW_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_cpasync<ROW_BYTES><<<n_cta, threads, smem>>>(tab, idx, rows_per_cta, valid_pct);
    CK(cudaEventRecord(e0));
    k_cpasync<ROW_BYTES><<<n_cta, threads, smem>>>(tab, idx, rows_per_cta, valid_pct);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    char m[64]; snprintf(m, 64, "cpasync t%d", threads); report(m, ms);
  }
  for (int threads : {256, 512, 1024}) {
    CK(cudaFuncSetAttribute(k_ldgsts<ROW_BYTES, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_ldgsts<ROW_BYTES, 8><<<n_cta, threads, smem>>>(tab, idx, rows_per_cta, valid_pct);
    CK(cudaEventRecord(e0));
    k_ldgsts<ROW_BYTES, 8><<<n_cta, threads, smem>>>(tab, idx, rows_per_cta, valid_pct);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    char m[64]; snprintf(m, 64, "ldg+sts t%d", threads); report(m, ms);
  }
  {
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    CUtensorMap map;
    cuuint64_t gd[2] = {(cuuint64_t)ROW_BYTES / 2, (cuuint64_t)n_rows_tab}; cuuint64_t gs[1] = {(cuuint64_t)ROW_BYTES};
    cuuint32_t box[2] = {(cuuint32_t)(ROW_BYTES > 128 ? 64 : ROW_BYTES / 2), 1}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fp)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, tab, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                ROW_BYTES >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed %d\n", (int)r);
    if (ROW_BYTES <= 128 && r == CUDA_SUCCESS) {
      CK(cudaFuncSetAttribute(k_tma_gather4<ROW_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_tma_gather4<ROW_BYTES><<<n_cta, 128, smem>>>(map, idx, rows_per_cta, valid_pct, n_rows_tab);
      CK(cudaEventRecord(e0));
      k_tma_gather4<ROW_BYTES><<<n_cta, 128, smem>>>(map, idx, rows_per_cta, valid_pct, n_rows_tab);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
      report("tma gather4", ms);
    }
  }
  {
    CK(cudaFuncSetAttribute(k_bulk_rows<ROW_BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bulk_rows<ROW_BYTES><<<n_cta, 128, smem>>>(tab, idx, rows_per_cta, valid_pct);
    CK(cudaEventRecord(e0));
    k_bulk_rows<ROW_BYTES><<<n_cta, 128, smem>>>(tab, idx, rows_per_cta, valid_pct);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    report("bulk rows (valid only)", ms);
  }
  CK(cudaFree(tab)); CK(cudaFree(idx));
}

int main() {
  const int rows_per_cta = 128 * 512;  // 64k rows per SM
  for (int valid : {100, 31}) {
    run<64>(2400000, rows_per_cta, valid, "154MB table (> L2)");
    run<128>(1200000, rows_per_cta, valid, "154MB table (> L2)");
    run<192>(800000, rows_per_cta, valid, "154MB table (> L2)");
    run<128>(150000, rows_per_cta, valid, "19MB table (L2 resident)");
  }
  return 0;
}
