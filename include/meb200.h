/*
 * meb200.h — C ABI of the B200-native sparse-convolution hot path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + the CUDA
 * stream to launch on (as `void*`, i.e. a cudaStream_t), allocates nothing the caller
 * did not hand in, and returns 0 on success or a negative MEB200_ERR_* code; the text
 * of the last failure on the calling thread is available from meb200_last_error().
 * No torch / C++ types appear in any signature.
 *
 * These are the calls the reference's FFI for this path would bind.  The reference's
 * boundary is the pybind module `MinkowskiEngineBackend._C` (pybind/extern.hpp); each
 * function below names the reference interface it stands in for.  The Python host in
 * `minkowskiengine_b200/` binds them through ctypes.
 *
 * Conventions
 *   coords      int32 [n, ncols] row-major, ncols = D+1, column 0 = batch index
 *               (reference: src/coordinate.hpp:36-61, MinkowskiSparseTensor.py:293-345)
 *   table       uint32 [capacity] open-addressing hash table of ROW INDICES into the
 *               coords array it was built over (0xFFFFFFFF = empty); capacity is a
 *               power of two from meb200_hash_capacity()
 *   nbr tables  int32 [K, n] "neighbour tables", k-major: nbr[k*n + r] = the row on
 *               the other side reached from row r through kernel offset k, or -1
 *   features    [n, C] row-major; dtype codes MEB200_F32 / MEB200_BF16 / MEB200_F16
 *   weights     [K, Cin, Cout] row-major (reference: MinkowskiConvolution.py:264-279)
 */
#ifndef MEB200_H_
#define MEB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEB200_OK 0
#define MEB200_ERR_INVALID (-1) /* bad argument (shape, dtype, null pointer)        */
#define MEB200_ERR_CUDA (-2)    /* a CUDA runtime call or kernel launch failed      */
#define MEB200_ERR_UNSUPPORTED (-3) /* valid request this build has no kernel for   */

#define MEB200_F32 0
#define MEB200_BF16 1
#define MEB200_F16 2

#define MEB200_MAX_NCOLS 8 /* D+1 <= 8, i.e. up to 7 spatial/temporal axes */

#define MEB200_POOL_SUM 0 /* reference PoolingMode::LOCAL_SUM_POOLING (src/types.hpp:141) */
#define MEB200_POOL_AVG 1 /* LOCAL_AVG_POOLING */
#define MEB200_POOL_MAX 2 /* LOCAL_MAX_POOLING */

/* ---- library ---------------------------------------------------------------------- */
const char *meb200_last_error(void);
/* Compile-time facts: "sm_100a", CUDA runtime version the library was built with.     */
const char *meb200_build_arch(void);
int meb200_cudart_version(void); /* reference: cudart_version(), pybind/extern.hpp:808-838 */
/* Number of kernels this library has launched since load (bench.py's gpu_launches).   */
uint64_t meb200_launch_count(void);
/* ... of which launches of the tcgen05 (tensor-core) convolution kernels. */
uint64_t meb200_tc_launch_count(void);

/* ---- coordinate hashing (reference a1/a2: src/coordinate.hpp:223-349,
 *      CoordinateMapCPU::insert_and_map coordinate_map_cpu.hpp:353-380,
 *      CoordinateMapGPU::insert coordinate_map_gpu.cu:196-278) ------------------------ */

/* Table capacity (power of two) used for n keys. */
uint32_t meb200_hash_capacity(uint32_t n);

/* Bytes of scratch the dedup pipeline needs for n candidate rows. */
uint64_t meb200_insert_scratch_bytes(uint32_t n);

/*
 * Deduplicating insert with the CPU reference's semantics: among equal coordinates the
 * FIRST row (smallest index) wins and unique rows are numbered by the rank of their
 * first occurrence.
 *   coords        [n, ncols] candidates          valid   optional uint8[n] (NULL = all)
 *   table         [capacity] out, built over `unique_coords` (row ids are NEW ids)
 *   unique_coords [n, ncols] out (first *h_num_unique rows are meaningful)
 *   unique_index  int64 [n] out: original row of each unique row (first m meaningful)
 *   inverse_map   int64 [n] out: new row id of every candidate (-1 where !valid)
 *   scratch       meb200_insert_scratch_bytes(n) bytes
 *   h_num_unique  HOST pointer; written after a stream synchronise (the one blocking
 *                 point of map construction: the caller needs m to size tensors).
 */
int meb200_insert_and_map(const int32_t *coords, const uint8_t *valid, uint32_t n,
                          uint32_t ncols, uint32_t *table, uint32_t capacity,
                          int32_t *unique_coords, int64_t *unique_index,
                          int64_t *inverse_map, void *scratch, uint32_t *h_num_unique,
                          void *stream);
/* The same insert WITHOUT the blocking count read, for maps that can be enqueued before they
 * are needed (the stride pyramid of a network: every level is a function of the input
 * coordinates alone, reference call sites coordinate_map_manager.cpp:406-429).  `n` is an UPPER
 * bound of the candidate count; `d_n` (device, may be NULL = all n) holds the actual count, so
 * a level can be enqueued while its parent's size is still only known on the device.  The unique
 * count is left in `d_num_unique` (DEVICE); `table` is dedup scratch here — build the real table
 * over the first m unique rows with meb200_map_build_table once m is known on the host.
 * Row numbering is identical to meb200_insert_and_map (rows past *d_n never win). */
int meb200_insert_and_map_enqueue(const int32_t *coords, const uint8_t *valid,
                                  const uint32_t *d_n, uint32_t n, uint32_t ncols,
                                  uint32_t *table, uint32_t capacity, int32_t *unique_coords,
                                  int64_t *unique_index, int64_t *inverse_map, void *scratch,
                                  uint32_t *d_num_unique, void *stream);
/* Row-index table over m DISTINCT coordinate rows (what meb200_insert_and_map leaves in `table`). */
int meb200_map_build_table(const int32_t *unique_coords, uint32_t m, uint32_t ncols,
                           uint32_t *table, uint32_t capacity, void *stream);

/* Candidate generation for a strided map: out[i] = floor(c / ts_out) * ts_out per
 * spatial axis, batch kept (reference a3: detail::stride_coordinate
 * src/coordinate_map.hpp:58-76 — integer floor division here).
 * out_tensor_stride: HOST int32[ncols-1]. */
int meb200_stride_coords(const int32_t *coords, uint32_t n, uint32_t ncols,
                         const int32_t *out_tensor_stride, int32_t *out, void *stream);

/* Candidate generation for stride_region (reference: CoordinateMapCPU::stride_region
 * coordinate_map_cpu.hpp:446-487): out[(i*K + k)] = coords[i] + offsets[k]; when
 * `aligned_only` != 0 candidates not aligned to out_tensor_stride get valid = 0.
 * offsets: DEVICE int32 [K, ncols-1]; out_tensor_stride: HOST int32[ncols-1]. */
int meb200_region_coords(const int32_t *coords, uint32_t n, uint32_t ncols,
                         const int32_t *offsets, uint32_t K,
                         const int32_t *out_tensor_stride, int aligned_only,
                         int32_t *out, uint8_t *valid, void *stream);

/* Batch lookup: result[i] = row of query[i] in the map, or -1
 * (reference: CoordinateMapCPU::find coordinate_map_cpu.hpp:388-412). */
int meb200_map_find(const int32_t *map_coords, const uint32_t *table, uint32_t capacity,
                    uint32_t ncols, const int32_t *query, uint32_t nq, int32_t *result,
                    void *stream);

/* ---- kernel map (reference a4/a5: kernel_region::coordinate_at kernel_region.hpp:198-247,
 *      CoordinateMapCPU::kernel_map coordinate_map_cpu.hpp:569-670,
 *      CoordinateMapGPU::kernel_map coordinate_map_gpu.cu:1549-1745) ------------------ */
/*
 * For every row x of the iterated map X and every offset k: probe map Y at
 * X[x] + offsets[k].  x_nbr[k*nx + x] = y (or -1) and, when y_nbr != NULL,
 * y_nbr[k*ny + y] = x (y_nbr must be pre-filled with -1 by the caller).
 * offsets: DEVICE int32 [K, ncols-1] already scaled by dilation * tensor stride.
 * d_num_pairs: optional DEVICE uint32 counter, incremented by the number of hits.
 */
int meb200_kernel_map(const int32_t *x_coords, uint32_t nx, const int32_t *y_coords,
                      uint32_t ny, const uint32_t *y_table, uint32_t y_capacity,
                      uint32_t ncols, const int32_t *offsets, uint32_t K, int32_t *x_nbr,
                      int32_t *y_nbr, uint32_t *d_num_pairs, void *stream);

/* Compacted pair lists of a neighbour table — the reference's own kernel-map representation
 * (gpu_kernel_map::in_maps / out_maps per offset, src/kernel_map.cuh:48-429; what
 * CoordinateMapManager::kernel_map returns, coordinate_map_manager.cpp:662-823).  The table rows
 * are cut into chunks of `chunk_rows` rows (rounded up to a multiple of 2048; 0 = one chunk);
 * segment s = chunk * K + k holds, in table-row order, pairs_other[i] = nbr[k][r], pairs_row[i] = r
 * for the rows r of the chunk with nbr[k][r] >= 0 and occupies [seg_start[s], seg_start[s+1]),
 * padded with (-1, -1) to a multiple of `stage` entries.  Deterministic, no atomics, no host
 * synchronisation.  Buffers: pairs_* hold meb200_pair_list_capacity(...) int32 each, seg_start
 * meb200_pair_list_chunks(...) * K + 1 int32, scratch meb200_pair_list_scratch_bytes(...) bytes. */
uint32_t meb200_pair_list_chunks(uint32_t n_rows, uint32_t chunk_rows);
uint64_t meb200_pair_list_scratch_bytes(uint32_t K, uint32_t n_rows, uint32_t chunk_rows);
uint64_t meb200_pair_list_capacity(uint32_t K, uint32_t n_rows, uint32_t stage, uint32_t chunk_rows);
int meb200_kernel_map_pairs(const int32_t *nbr, uint32_t K, uint32_t n_rows, uint32_t stage,
                            uint32_t chunk_rows, int32_t *pairs_other, int32_t *pairs_row,
                            int32_t *seg_start, void *scratch, void *stream);

/* ---- sparse convolution (reference a8/a9/a10: ConvolutionForwardKernelCPU /
 *      ConvolutionBackwardKernelCPU src/convolution_kernel.hpp:33-144, GPU
 *      src/convolution_kernel.cu:114-496,553-757) ------------------------------------ */
/*
 * out[o,:] = sum_k in[out_nbr[k,o],:] @ W[k]          (rows with out_nbr = -1 skipped)
 * Output is fully written (no pre-zeroing needed, no atomics, deterministic).
 * in_dtype: dtype of `in` and `weight`; out_dtype: dtype of `out` (F32 always allowed).
 */
int meb200_conv_forward(const void *in, int in_dtype, uint32_t n_in, uint32_t c_in,
                        const void *weight, uint32_t K, uint32_t c_out,
                        const int32_t *out_nbr, uint32_t n_out, void *out, int out_dtype,
                        void *workspace, uint64_t workspace_bytes, void *stream);

/* grad_in[i,:] = sum_k grad_out[in_nbr[k,i],:] @ W[k]^T   (fully written)
 * grad_weight[k] = sum_o in[out_nbr[k,o],:]^T @ grad_out[o,:]   (fully written, fp32) */
int meb200_conv_backward(const void *in, const void *grad_out, int dtype, uint32_t n_in,
                         uint32_t c_in, const void *weight, uint32_t K, uint32_t c_out,
                         const int32_t *out_nbr, const int32_t *in_nbr, uint32_t n_out,
                         void *grad_in, int grad_in_dtype, float *grad_weight,
                         void *workspace, uint64_t workspace_bytes, void *stream);

/* Weight packing for the tensor-core path, once per optimizer step instead of once per call:
 * fp32 master weight [K, Cin, Cout] -> `dtype` (BF16/F16) operand copies
 *   w_cast [K, Cin, Cout]                 w_t  [K, Cout, Cin]
 *   w_cp   [K, Cin, perm(Cout)]           w_tp [K, Cout, perm(Cin)]
 * where perm reorders the reduction axis inside every block of 32 channels to the order in
 * which the A-in-tensor-memory kernel lays gathered rows out (position 16h+4j+2e+b holds
 * channel 8j+4h+2e+b).  w_cp / w_tp may be NULL and are skipped when Cout / Cin is not a
 * multiple of 32.  The reference keeps weights in the feature dtype and needs no such step
 * (MinkowskiConvolution.py:264-279); this is the bf16/fp16 operand cache of this backend. */
int meb200_conv_pack_weights(const float *weight, uint32_t K, uint32_t c_in, uint32_t c_out,
                             int dtype, void *w_cast, void *w_t, void *w_cp, void *w_tp,
                             void *stream);

/* The same packing for MANY layers in one launch (all weights of a network change together, at
 * the optimizer step): `jobs_dev` is a DEVICE array of n_jobs jobs; job j owns the 32x32 tiles
 * [tile_begin_j, tile_begin_j + K_j * ceil(c_in_j / 32) * ceil(c_out_j / 32)) of the launch,
 * tile_begin ascending from 0, total_tiles = their sum.  The table only has to be rebuilt when
 * the set of tensors (or their addresses) changes. */
typedef struct meb200_pack_job {
  const void *w;                      /* fp32 [K, c_in, c_out] */
  void *w_cast, *w_t, *w_cp, *w_tp;   /* as meb200_conv_pack_weights; w_cp / w_tp may be NULL */
  uint32_t K, c_in, c_out;
  uint32_t tile_begin;
} meb200_pack_job;
int meb200_conv_pack_weights_batched(const meb200_pack_job *jobs_dev, uint32_t n_jobs,
                                     uint32_t total_tiles, int dtype, void *stream);

/* Network stem (the first layer of a network: c_in <= 4 input channels, e.g. RGB; reference
 * call site examples/minkunet.py:113-116 conv0p1s1, kernel 5 -> K = 125).  Rows of 8 bytes defeat
 * the 64-byte-block gather of the general kernels, so the layer is run as a K = 1 convolution over
 * VIRTUAL channels v = 4 k + c (offset k < 16 ceil(K / 16), channel c < 4):
 *   in4            [n_in, 4]   the features zero-padded to 4 channels
 *   weight_v       [c_out, V]  V = meb200_conv_stem_virtual_channels(K); the `w_tp` output of
 *                              meb200_conv_pack_weights(Wv, K=1, c_in=V, c_out) where
 *                              Wv[4 k + c][n] = W[k][c][n] (zero for padded k, c)
 *   grad_weight_v  [V, c_out]  fp32, same indexing (zero-filled by the call)
 * bf16 / fp16 only; c_out a multiple of 16, <= 256; K <= 128. */
uint32_t meb200_conv_stem_virtual_channels(uint32_t K);
int meb200_conv_stem_supported(int dtype, uint32_t K, uint32_t c_out);
int meb200_conv_stem_forward(const void *in4, int dtype, uint32_t K, const void *weight_v,
                             uint32_t c_out, const int32_t *out_nbr, uint32_t n_out, void *out,
                             int out_dtype, void *stream);
int meb200_conv_stem_wgrad(const void *in4, const void *grad_out, int dtype, uint32_t K,
                           uint32_t c_out, const int32_t *out_nbr, uint32_t n_out,
                           float *grad_weight_v, void *stream);

/* meb200_conv_forward with the weight already packed (tensor-core path only: returns
 * MEB200_ERR_UNSUPPORTED for shapes/dtypes outside it, the caller then uses
 * meb200_conv_forward).  Same result, no per-call cast/transpose, no workspace.
 * weight_tp (may be NULL) selects the A-in-tensor-memory kernel when Cin % 32 == 0. */
int meb200_conv_forward_packed(const void *in, int in_dtype, uint32_t n_in, uint32_t c_in,
                               const void *weight_t, const void *weight_tp, uint32_t K,
                               uint32_t c_out, const int32_t *out_nbr, uint32_t n_out, void *out,
                               int out_dtype, void *stream);

/* meb200_conv_backward on packed weights (w_cast, and w_cp or NULL), tensor-core path only;
 * grad_in / grad_weight may be NULL to skip dgrad / wgrad.  pairs_in / pairs_out / seg_start
 * (all three or NULL) + n_chunks: the compacted pair lists of out_nbr from
 * meb200_kernel_map_pairs with stage = 64 and their chunk count; with them wgrad reduces over
 * valid pairs only instead of the dense table. */
int meb200_conv_backward_packed(const void *in, const void *grad_out, int dtype, uint32_t n_in,
                                uint32_t c_in, const void *w_cast, const void *w_cp, uint32_t K,
                                uint32_t c_out, const int32_t *out_nbr, const int32_t *in_nbr,
                                uint32_t n_out, void *grad_in, int grad_in_dtype,
                                float *grad_weight, const int32_t *pairs_in,
                                const int32_t *pairs_out, const int32_t *seg_start,
                                uint32_t n_chunks, void *stream);

/* Workspace the two calls above may use (0 is legal: slower fallbacks are chosen). */
uint64_t meb200_conv_workspace_bytes(uint32_t n_in, uint32_t n_out, uint32_t c_in,
                                     uint32_t c_out, uint32_t K, int dtype);

/* ---- local pooling (reference a11: src/pooling_avg_kernel.hpp:40-150,
 *      src/pooling_max_kernel.hpp:35-115, src/local_pooling_cpu.cpp:43-185) ---------- */
/* aux: AVG -> num_nonzero [n_out] (feature dtype); MAX -> max_index int32 [n_out, C]
 * (flat index row*C + c into `in`, -1 for rows without input); SUM -> unused (NULL ok). */
int meb200_pool_forward(const void *in, int dtype, uint32_t n_in, uint32_t C,
                        const int32_t *out_nbr, uint32_t K, uint32_t n_out, int mode,
                        void *out, void *aux, void *stream);
/* grad_in fully written for SUM/AVG; for MAX the caller zero-fills grad_in first. */
int meb200_pool_backward(const void *grad_out, int dtype, uint32_t n_in, uint32_t C,
                         const int32_t *in_nbr, uint32_t K, uint32_t n_out, int mode,
                         const void *aux, void *grad_in, void *stream);

/* ---- batch normalisation over [n, C] feature rows (SURVEY.md 8(f) "next" row 1; the
 *      reference applies torch.nn.BatchNorm1d to `.F`: MinkowskiEngine/MinkowskiNormalization.py:51-99).
 *      C must be a multiple of 8.  sums: DEVICE double [2C]; statistics tensors fp32 [C]. ------ */
/* mean / invstd (biased variance + eps) from the (possibly all-reduced) sums
 * [sum x | sum x^2] over `count` rows; running statistics (may be NULL) updated with momentum
 * and the unbiased variance.  The row count is `count`, or — for synchronised BN over NCCL,
 * where it is the all-reduced total that only exists on the device — read from *d_count when
 * d_count != NULL. */
int meb200_bn_finalize(const double *sums, double count, const double *d_count, uint32_t C,
                       float eps, float momentum, float *running_mean, float *running_var,
                       float *mean, float *invstd, void *stream);
/* The streaming passes, with the BasicBlock tail of modules/resnet_block.py:52-68 folded in:
 *   apply_fused           y = relu?( bn(x) + residual? ),  bn(x) = (x - mean) * invstd * weight + bias
 *   backward_apply_fused  with dy' = relu ? dy * (y_mask > 0) : dy  (y_mask = the fused output y):
 *                         dx = (dy' - sums[c]/count - xhat * sums[C+c]/count) * invstd * weight,
 *                         d_residual (may be NULL) = dy'
 * weight / bias / residual / y_mask / d_residual may be NULL. */
int meb200_bn_apply_fused(const void *x, int dtype, uint32_t n, uint32_t C, const float *mean,
                          const float *invstd, const float *weight, const float *bias,
                          const void *residual, int relu, void *y, void *stream);
int meb200_bn_backward_apply_fused(const void *dy, const void *x, const void *y_mask, int dtype,
                                   uint32_t n, uint32_t C, const float *mean, const float *invstd,
                                   const float *weight, const double *sums, double count,
                                   const double *d_count, void *dx, void *d_residual,
                                   void *stream);

/* The reductions (two launches per pass, no memset): they take a WORKSPACE of
 * meb200_bn_workspace_bytes() bytes, zero-filled once by the caller and shared by every layer
 * that runs on the same stream; the last CTA of a reduction consumes the totals (finalize /
 * parameter gradients / copy-out) and leaves the workspace zero again.
 *   meb200_bn_forward_train      = stats + finalize (one launch) + apply_fused; n > 0;
 *                                  *num_batches_tracked (int64, may be NULL) is incremented
 *   meb200_bn_stats_to           = stats, totals [sum x | sum x^2] copied to sums_out[2C]
 *                                  (the exchange slot of a synchronised layer)
 *   meb200_bn_backward_reduce_to = backward reduce, totals [sum dy | sum dy*xhat] copied to
 *                                  sums_out[2C] and, when non-NULL, to fp32 grad_bias / grad_weight */
uint64_t meb200_bn_workspace_bytes(void);
int meb200_bn_forward_train(const void *x, int dtype, uint32_t n, uint32_t C, const float *weight,
                            const float *bias, const void *residual, int relu, float eps,
                            float momentum, float *running_mean, float *running_var,
                            long long *num_batches_tracked, void *workspace, float *mean,
                            float *invstd, void *y, void *stream);
int meb200_bn_stats_to(const void *x, int dtype, uint32_t n, uint32_t C, void *workspace,
                       double *sums_out, void *stream);
int meb200_bn_backward_reduce_to(const void *dy, const void *x, const void *y_mask, int dtype,
                                 uint32_t n, uint32_t C, const float *mean, const float *invstd,
                                 void *workspace, double *sums_out, float *grad_weight,
                                 float *grad_bias, void *stream);

/* ---- synchronised batch norm: statistics exchange over NVLink peer memory ----------------
 * Replaces the per-layer NCCL all-reduce of torch.nn.SyncBatchNorm (reference:
 * MinkowskiEngine/MinkowskiNormalization.py:101-192, examples/multigpu_ddp.py:91-95).  Every rank
 * owns a buffer in symmetric memory with the same layout: bytes [0, 1024) hold uint32 flags
 * (flag[r] = last call whose slot of rank r is complete), slots follow.  `peer_bases_dev` is a
 * DEVICE array of `world` device pointers, entry r = base of rank r's buffer as mapped into this
 * process.  An exchange publishes `seq` to every peer, waits for every peer's `seq`, then sums
 * the ranks' slots in rank order (bitwise identical on all ranks).  `seq` must increase by one
 * per call on every rank; consecutive calls must use different slots (a rank can be at most one
 * call ahead of its slowest peer).  Flags must be zero before the first call.
 * The exchange is run by the LAST CTA of the batch-norm reductions themselves (no separate
 * launch): meb200_bn_forward_train / meb200_bn_backward_reduce_to of a synchronised layer.  The
 * slot ([sum | sum2 | rows] forward, [sum dy | sum dy*xhat] backward) is filled by the kernel;
 * mean / 1/std / running statistics come from the global totals and the global row count
 * (returned in *total_rows, device); the parameter gradients stay LOCAL sums (DDP averages them),
 * sums_out receives the global ones for meb200_bn_backward_apply_fused.  A rank with n = 0 rows
 * still takes part. */
int meb200_bn_forward_train_peer(const void *x, int dtype, uint32_t n, uint32_t C,
                                 const float *weight, const float *bias, const void *residual,
                                 int relu, float eps, float momentum, float *running_mean,
                                 float *running_var, long long *num_batches_tracked,
                                 void *workspace, const void *peer_bases_dev,
                                 uint64_t slot_offset_bytes, uint32_t seq, uint32_t rank,
                                 uint32_t world, float *mean, float *invstd, double *total_rows,
                                 void *y, void *stream);
int meb200_bn_backward_reduce_peer(const void *dy, const void *x, const void *y_mask, int dtype,
                                   uint32_t n, uint32_t C, const float *mean, const float *invstd,
                                   void *workspace, const void *peer_bases_dev,
                                   uint64_t slot_offset_bytes, uint32_t seq, uint32_t rank,
                                   uint32_t world, double *sums_out, float *grad_weight,
                                   float *grad_bias, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MEB200_H_ */
