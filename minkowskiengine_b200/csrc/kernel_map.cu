// Kernel-map construction: per-offset in/out row pairs as k-major neighbour tables.
//
// Replaces the reference's direct_kernel_map + remove_if + sort_by_key decomposition
// (src/coordinate_map_gpu.cu:1479-1542,1697-1733, src/kernel_map.cuh:313-405) and the CPU
// loop it mirrors (src/coordinate_map_cpu.hpp:569-670).  One thread per (offset k, row x)
// with x fastest: coordinate loads and table writes are coalesced, each probe is one
// hash-table sector + one coordinate sector.  The result is written directly in the
// layout the convolution kernels consume — x_nbr[k][x] (stationary side) and
// y_nbr[k][y] (its transpose, used by dgrad / transposed convolution) — so there is no
// compaction, no sort and no host synchronisation on this path, and the pair order is
// deterministic.  The reference's per-offset pair lists are recoverable as the
// non-negative entries of row k (that is what CoordinateManager.kernel_map() returns).
#include "common.cuh"

namespace meb200 {

template <int NC>
__global__ void __launch_bounds__(256)
k_kernel_map(const int32_t *__restrict__ x_coords, uint32_t nx,
             const int32_t *__restrict__ y_coords, uint32_t ny,
             const uint32_t *__restrict__ y_table, uint32_t mask,
             const int32_t *__restrict__ offsets, uint32_t K, int32_t *__restrict__ x_nbr,
             int32_t *__restrict__ y_nbr, uint32_t *__restrict__ num_pairs) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool active = t < (uint64_t)nx * K;
  int32_t y = -1;
  if (active) {
    uint32_t k = (uint32_t)(t / nx), x = (uint32_t)(t % nx);
    int32_t key[NC];
    load_coord<NC>(x_coords, x, key);
#pragma unroll
    for (int j = 1; j < NC; ++j) key[j] += __ldg(offsets + (size_t)k * (NC - 1) + (j - 1));
    y = table_find<NC>(y_coords, y_table, mask, key);
    x_nbr[t] = y;
    if (y >= 0 && y_nbr != nullptr) y_nbr[(size_t)k * ny + y] = (int32_t)x;
  }
  if (num_pairs != nullptr) {
    unsigned hits = __popc(__ballot_sync(0xffffffffu, y >= 0));
    if ((threadIdx.x & 31) == 0 && hits) atomicAdd(num_pairs, hits);
  }
}

}  // namespace meb200

using namespace meb200;

extern "C" int meb200_kernel_map(const int32_t *x_coords, uint32_t nx, const int32_t *y_coords,
                                 uint32_t ny, const uint32_t *y_table, uint32_t y_capacity,
                                 uint32_t ncols, const int32_t *offsets, uint32_t K,
                                 int32_t *x_nbr, int32_t *y_nbr, uint32_t *d_num_pairs,
                                 void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  MEB_CHECK_ARG(y_capacity >= 2 && (y_capacity & (y_capacity - 1)) == 0, "capacity=%u",
                y_capacity);
  if (nx == 0 || K == 0) return MEB200_OK;
  MEB_CHECK_ARG(x_coords && y_table && offsets && x_nbr, "null buffer");
  MEB_CHECK_ARG(ny == 0 || y_coords != nullptr, "y_coords");
  uint64_t total = (uint64_t)nx * K;
  MEB_CHECK_ARG(total < (1ull << 40), "nx*K too large");
  MEB_DISPATCH_NCOLS(ncols, k_kernel_map<NC><<<cdiv(total, 256), 256, 0, stream>>>(
                                x_coords, nx, y_coords, ny, y_table, y_capacity - 1, offsets, K,
                                x_nbr, y_nbr, d_num_pairs));
  MEB_LAUNCH_OK();
  return MEB200_OK;
}
