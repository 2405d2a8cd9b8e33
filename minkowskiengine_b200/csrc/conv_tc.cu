// tcgen05 sparse convolution: output-stationary implicit GEMM over the k-major neighbour
// table, bf16/fp16 operands, fp32 accumulation in tensor memory.
//
//   out[r, :] = sum_k  A[nbr[k][r], :] @ B_k          r in a 128-row tile
//
// Replaces the reference's per-offset SIMT tile-matmul with per-element atomicAdd
// (src/convolution_kernel.cu:114-180,320-496): ONE persistent launch per layer, every
// output row written exactly once (no atomics, no zero-fill, deterministic), the fp32
// accumulator never leaves the SM until the epilogue converts it.
//
// CTA = 9 warps, one CTA per SM, persistent over "super tiles" of R x 128 output rows:
//   warps 0-3  producers : gather A rows (cp.async 16 B, zero-fill for missing neighbours)
//                          into 128B/64B/32B-swizzled K-major tiles; load the B_k slice once
//                          per (offset, channel chunk) and reuse it for the R row tiles
//   warps 4-7  epilogue  : tcgen05.ld accumulator -> convert -> 16 B global stores
//   warp  8    MMA       : lane 0 issues tcgen05.mma (M=128, N=c_cols, K=16) and the
//                          tcgen05.commit that recycles smem stages / publishes accumulators
// Pipelines (all mbarrier based): A ring (nA stages), B ring (nB stages), accumulator
// ring in TMEM (acc_stages x R accumulators of c_cols columns).
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace meb200 {

using namespace ptx;

constexpr int kTileM = 128;
constexpr int kProducerWarps = 4;
constexpr int kEpilogueWarps = 4;
constexpr int kThreads = (kProducerWarps + kEpilogueWarps + 1) * 32;
constexpr int kMaxStages = 12;
constexpr int kLag = 3;  // cp.async groups kept in flight per producer thread
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSmemBudget = 200 * 1024;

struct TcParams {
  const void *A;        // [n_a, c_red] gathered operand
  const void *Wb;       // [K, c_cols, c_red] operand B, reduction dim contiguous
  const int32_t *nbr;   // [K, n_rows]
  void *out;            // [n_rows, c_cols]
  uint32_t c_red, c_cols, K, n_rows;
  uint32_t n_super, R, acc_stages, nA, nB;
  uint32_t a_stage_bytes, b_stage_bytes;
  uint32_t out_f32;
};

template <int BK>
struct Swz {
  static constexpr int kRowBytes = BK * 2;
  static constexpr int kChunks = kRowBytes / 16;          // 16-byte chunks per row
  static constexpr int kRowsPerInstr = 32 / kChunks;      // rows one warp-wide cp.async covers
  static constexpr int kRowsPerIter = kRowsPerInstr * kProducerWarps;
  static constexpr int kItersA = kTileM / kRowsPerIter;
  static constexpr uint32_t kLayout = BK == 64 ? 2u : (BK == 32 ? 4u : 6u);
  static constexpr uint32_t kSbo = 8 * kRowBytes;         // 8-row core-matrix group stride
  __device__ static __forceinline__ uint32_t offset(uint32_t row, uint32_t chunk) {
    uint32_t x = BK == 64 ? (row & 7u) : (BK == 32 ? ((row >> 1) & 3u) : ((row >> 2) & 1u));
    return row * kRowBytes + ((chunk ^ x) << 4);
  }
};

template <typename T> struct Fmt;
template <> struct Fmt<__nv_bfloat16> { static constexpr uint32_t v = 1; };
template <> struct Fmt<__half> { static constexpr uint32_t v = 0; };

template <typename T>
__device__ __forceinline__ uint32_t pack2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}

template <typename T, int BK>
__global__ void __launch_bounds__(kThreads, 1) k_conv_tc(const TcParams p) {
  using S = Swz<BK>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16 B aligned: round up to the swizzle atom
  uint8_t *smem = reinterpret_cast<uint8_t *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + p.nA * p.a_stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.nA * p.a_stage_bytes +
                                                p.nB * p.b_stage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  // barrier slots: fullA[kMax] emptyA[kMax] fullB[kMax] emptyB[kMax] tfull[2] tempty[2]
  auto fullA = [&](uint32_t i) { return bar0 + 8u * i; };
  auto emptyA = [&](uint32_t i) { return bar0 + 8u * (kMaxStages + i); };
  auto fullB = [&](uint32_t i) { return bar0 + 8u * (2 * kMaxStages + i); };
  auto emptyB = [&](uint32_t i) { return bar0 + 8u * (3 * kMaxStages + i); };
  auto tfull = [&](uint32_t i) { return bar0 + 8u * (4 * kMaxStages + i); };
  auto tempty = [&](uint32_t i) { return bar0 + 8u * (4 * kMaxStages + 2 + i); };
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 * kMaxStages + 4);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t chunks_red = p.c_red / BK;
  const uint32_t R = p.R;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < p.nA; ++i) {
      mbar_init(fullA(i), kProducerWarps * 32);
      mbar_init(emptyA(i), 1);
    }
    for (uint32_t i = 0; i < p.nB; ++i) {
      mbar_init(fullB(i), kProducerWarps * 32);
      mbar_init(emptyB(i), 1);
    }
    for (uint32_t i = 0; i < 2; ++i) {
      mbar_init(tfull(i), 1);
      mbar_init(tempty(i), kEpilogueWarps * 32);
    }
    mbar_fence_init();
  }
  if (warp == kProducerWarps + kEpilogueWarps) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kProducerWarps) {
    // =============================== PRODUCERS ======================================
    const uint32_t sub_row = lane / S::kChunks, chunk = lane % S::kChunks;
    const T *A = reinterpret_cast<const T *>(p.A);
    const T *Wb = reinterpret_cast<const T *>(p.Wb);
    const uint32_t iters_b = (p.c_cols + S::kRowsPerIter - 1) / S::kRowsPerIter;
    uint32_t g = 0;         // A-stage groups committed so far (one cp.async group each)
    uint32_t signaled = 0;  // groups already published to the MMA warp

    auto signal = [&](uint32_t gd) {  // group gd has landed in smem
      fence_proxy_async();
      mbar_arrive(fullA(gd % p.nA));
      if (gd % R == 0) mbar_arrive(fullB((gd / R) % p.nB));
    };
    // Wait for a stage to be recycled.  If the consumer is behind, first publish everything
    // this thread still holds back (the consumer may need exactly those groups to progress).
    auto wait_empty = [&](uint32_t bar, uint32_t parity) {
      if (mbar_try_wait(bar, parity)) return;
      cp_async_wait<0>();
      while (signaled < g) signal(signaled++);
      mbar_wait(bar, parity);
    };

    for (uint32_t st = blockIdx.x; st < p.n_super; st += gridDim.x) {
      const uint32_t row_base = st * R * kTileM;
      for (uint32_t k = 0; k < p.K; ++k) {
        const int32_t *nbr_k = p.nbr + (size_t)k * p.n_rows;
        for (uint32_t cc = 0; cc < chunks_red; ++cc) {
          const uint32_t c0 = cc * BK;
          for (uint32_t r = 0; r < R; ++r) {
            const uint32_t sa = g % p.nA;
            wait_empty(emptyA(sa), ((g / p.nA) & 1u) ^ 1u);
            if (r == 0) {
              const uint32_t gb = g / R, sb = gb % p.nB;
              wait_empty(emptyB(sb), ((gb / p.nB) & 1u) ^ 1u);
              const uint32_t dstB = sB + sb * p.b_stage_bytes;
              for (uint32_t j = 0; j < iters_b; ++j) {
                uint32_t n = j * S::kRowsPerIter + warp * S::kRowsPerInstr + sub_row;
                if (n < p.c_cols) {
                  const T *src = Wb + ((size_t)k * p.c_cols + n) * p.c_red + c0 + chunk * 8;
                  cp_async16(dstB + S::offset(n, chunk), src, 16);
                }
              }
            }
            const uint32_t dstA = sA + sa * p.a_stage_bytes;
            const uint32_t tile_row0 = row_base + r * kTileM;
#pragma unroll
            for (int j = 0; j < S::kItersA; ++j) {
              uint32_t row = j * S::kRowsPerIter + warp * S::kRowsPerInstr + sub_row;
              uint32_t grow = tile_row0 + row;
              int32_t src_row = (grow < p.n_rows) ? __ldg(nbr_k + grow) : -1;
              const T *src = A + (size_t)(src_row < 0 ? 0 : src_row) * p.c_red + c0 + chunk * 8;
              cp_async16(dstA + S::offset(row, chunk), src, src_row < 0 ? 0u : 16u);
            }
            cp_async_commit();
            ++g;
            if (g - signaled > (uint32_t)kLag) {
              cp_async_wait<kLag>();          // all but the newest kLag groups have landed
              while (g - signaled > (uint32_t)kLag) signal(signaled++);
            }
          }
        }
      }
    }
    cp_async_wait<0>();
    while (signaled < g) signal(signaled++);
  } else if (warp < kProducerWarps + kEpilogueWarps) {
    // =============================== EPILOGUE =======================================
    const uint32_t q = warp & 3u;  // TMEM lane quadrant this warp may access
    uint32_t t = 0;
    for (uint32_t st = blockIdx.x; st < p.n_super; st += gridDim.x, ++t) {
      const uint32_t acc = t % p.acc_stages, par = (t / p.acc_stages) & 1u;
      mbar_wait(tfull(acc), par);
      tc_fence_after();
      for (uint32_t r = 0; r < R; ++r) {
        const uint32_t row = (st * R + r) * kTileM + q * 32 + lane;
        const uint32_t col0 = (acc * R + r) * p.c_cols;
        for (uint32_t c = 0; c < p.c_cols; c += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((q * 32u) << 16) + col0 + c, v);
          tmem_ld_wait();
          if (row < p.n_rows) {
            if (p.out_f32) {
              float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out) +
                                                       (size_t)row * p.c_cols + c);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                     __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            } else {
              uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<T *>(p.out) +
                                                     (size_t)row * p.c_cols + c);
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                uint4 w;
                w.x = pack2<T>(__uint_as_float(v[8 * i + 0]), __uint_as_float(v[8 * i + 1]));
                w.y = pack2<T>(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3]));
                w.z = pack2<T>(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5]));
                w.w = pack2<T>(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7]));
                dst[i] = w;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty(acc));
    }
  } else if (lane == 0) {
    // =============================== MMA ISSUER =====================================
    const uint32_t idesc = umma_idesc_f16(Fmt<T>::v, kTileM, p.c_cols, 0, 0);
    uint32_t g = 0, t = 0;
    for (uint32_t st = blockIdx.x; st < p.n_super; st += gridDim.x, ++t) {
      const uint32_t acc = t % p.acc_stages, par = (t / p.acc_stages) & 1u;
      mbar_wait(tempty(acc), par ^ 1u);
      tc_fence_after();
      for (uint32_t k = 0; k < p.K; ++k) {
        for (uint32_t cc = 0; cc < chunks_red; ++cc) {
          const uint32_t gb = g / R, sb = gb % p.nB;
          mbar_wait(fullB(sb), (gb / p.nB) & 1u);
          const uint32_t bbase = sB + sb * p.b_stage_bytes;
          for (uint32_t r = 0; r < R; ++r, ++g) {
            const uint32_t sa = g % p.nA;
            mbar_wait(fullA(sa), (g / p.nA) & 1u);
            tc_fence_after();
            const uint32_t abase = sA + sa * p.a_stage_bytes;
            const uint32_t d = tmem_base + (acc * R + r) * p.c_cols;
#pragma unroll
            for (int j = 0; j < BK / 16; ++j) {
              uint64_t da = umma_desc(abase + j * 32, 16, S::kSbo, S::kLayout);
              uint64_t db = umma_desc(bbase + j * 32, 16, S::kSbo, S::kLayout);
              umma_f16(d, da, db, idesc, (k | cc | (uint32_t)j) != 0u);
            }
            umma_commit(emptyA(sa));
          }
          umma_commit(emptyB(sb));
        }
      }
      umma_commit(tfull(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarps + kEpilogueWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// Wb[k][n][c] = W[k][c][n]  (forward operand B: reduction dim contiguous)
template <typename T>
__global__ void __launch_bounds__(256)
k_transpose_w(const T *__restrict__ W, T *__restrict__ Wb, uint32_t c_in, uint32_t c_out) {
  __shared__ T tile[32][33];
  const T *Wk = W + (size_t)blockIdx.z * c_in * c_out;
  T *Wbk = Wb + (size_t)blockIdx.z * c_in * c_out;
  uint32_t ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (uint32_t i = ty; i < 32; i += 8)
    if (ci0 + i < c_in && co0 + tx < c_out) tile[i][tx] = Wk[(size_t)(ci0 + i) * c_out + co0 + tx];
  __syncthreads();
  for (uint32_t i = ty; i < 32; i += 8)
    if (co0 + i < c_out && ci0 + tx < c_in) Wbk[(size_t)(co0 + i) * c_in + ci0 + tx] = tile[tx][i];
}

static int pick_bk(uint32_t c_red) {
  if (c_red % 64 == 0) return 64;
  if (c_red % 32 == 0) return 32;
  if (c_red % 16 == 0) return 16;
  return 0;
}

bool conv_tc_supported(int dtype, uint32_t c_reduce, uint32_t c_cols) {
  if (dtype != MEB200_BF16 && dtype != MEB200_F16) return false;
  if (pick_bk(c_reduce) == 0) return false;
  return c_cols % 16 == 0 && c_cols >= 16 && c_cols <= 256;
}

template <typename T, int BK>
static int launch_tc(const TcParams &p, size_t smem_bytes, unsigned grid, cudaStream_t stream) {
  auto kern = k_conv_tc<T, BK>;
  static bool configured = false;
  if (!configured) {
    MEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  kern<<<grid, kThreads, smem_bytes, stream>>>(p);
  count_tc_launch();
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

int conv_forward_tc(const void *A, int dtype, uint32_t n_a, uint32_t c_reduce, const void *W,
                    uint32_t K, uint32_t c_cols, bool dgrad, const int32_t *nbr, uint32_t n_rows,
                    void *out, int out_dtype, void *workspace, cudaStream_t stream) {
  (void)n_a;
  if (n_rows == 0) return MEB200_OK;
  const int bk = pick_bk(c_reduce);
  MEB_CHECK_ARG(bk != 0 && c_cols % 16 == 0 && c_cols <= 256, "shape not supported by tc path");
  const void *Wb = W;
  if (!dgrad) {
    // forward: B_k = W[k]^T so that the reduction dim (c_in) is contiguous
    MEB_CHECK_ARG(workspace != nullptr, "workspace required");
    dim3 grid(cdiv(c_cols, 32), cdiv(c_reduce, 32), K);
    if (dtype == MEB200_BF16)
      k_transpose_w<__nv_bfloat16><<<grid, 256, 0, stream>>>(
          (const __nv_bfloat16 *)W, (__nv_bfloat16 *)workspace, c_reduce, c_cols);
    else
      k_transpose_w<__half><<<grid, 256, 0, stream>>>((const __half *)W, (__half *)workspace,
                                                     c_reduce, c_cols);
    MEB_LAUNCH_OK();
    Wb = workspace;
  }
  TcParams p{};
  p.A = A; p.Wb = Wb; p.nbr = nbr; p.out = out;
  p.c_red = c_reduce; p.c_cols = c_cols; p.K = K; p.n_rows = n_rows;
  p.out_f32 = out_dtype == MEB200_F32;
  // accumulators: R row tiles share one B slice; double-buffer the set when it fits in TMEM
  uint32_t R = kTmemCols / (2 * c_cols);
  if (R > 4) R = 4;
  uint32_t acc_stages = 2;
  if (R == 0) { R = 1; acc_stages = kTmemCols / c_cols >= 2 ? 2 : 1; }
  uint32_t tiles = cdiv(n_rows, kTileM);
  if (R > tiles) R = tiles;
  p.R = R; p.acc_stages = acc_stages;
  p.n_super = cdiv(tiles, R);
  p.a_stage_bytes = kTileM * bk * 2;
  p.b_stage_bytes = ((c_cols * bk * 2 + 1023) / 1024) * 1024;
  p.nB = 3;
  const uint32_t bar_bytes = (4 * kMaxStages + 4) * 8 + 16;
  uint32_t avail = kSmemBudget - p.nB * p.b_stage_bytes - bar_bytes - 1024;
  p.nA = avail / p.a_stage_bytes;
  if (p.nA > (uint32_t)kMaxStages) p.nA = kMaxStages;
  MEB_CHECK_ARG(p.nA >= 4, "not enough shared memory for the A pipeline");
  size_t smem_bytes = 1024 + (size_t)p.nA * p.a_stage_bytes + (size_t)p.nB * p.b_stage_bytes + bar_bytes;
  unsigned grid = p.n_super < (uint32_t)num_sms() ? p.n_super : (unsigned)num_sms();
#define MEB_TC_LAUNCH(TT)                                                          \
  switch (bk) {                                                                    \
    case 64: return launch_tc<TT, 64>(p, smem_bytes, grid, stream);                \
    case 32: return launch_tc<TT, 32>(p, smem_bytes, grid, stream);                \
    default: return launch_tc<TT, 16>(p, smem_bytes, grid, stream);                \
  }
  if (dtype == MEB200_BF16) { MEB_TC_LAUNCH(__nv_bfloat16) }
  MEB_TC_LAUNCH(__half)
#undef MEB_TC_LAUNCH
}

// =====================================================================================
// wgrad:  dW[k][ci][co] = sum_o In[nbr[k][o]][ci] * dOut[o][co]
//
// GEMM view: D[M = ci][N = co] += A[M][Kd] * B[Kd][N] with the reduction dim Kd = output rows.
// Both operands are gathered/loaded as ROWS (one row = the channels of one voxel), i.e. they sit
// in shared memory with the M / N dim contiguous: "MN-major" UMMA operands.  Physical layout per
// 64-channel block: [rows][128 B] with the same 8-row / 128B-swizzle atoms as the forward tiles
// (so the gather code is shared); LBO = distance between 64-channel blocks, SBO = 1024 B between
// 8-row groups.  One CTA owns (a group of G offsets, a slice of the output rows, <= 2 m-tiles of
// 128 input channels); the dOut chunk (operand B) is loaded once and reused for the G offsets,
// each offset accumulating into its own TMEM accumulator; the epilogue adds the fp32 partials
// into dW with red.global (dW is zero-filled first).
// =====================================================================================
constexpr int kWgRows = 64;                 // reduction rows per pipeline stage
constexpr uint32_t kBlkBytes = kWgRows * 128;  // one 64-channel block of one stage

struct WgParams {
  const void *in;        // [n_in, c_in]
  const void *gout;      // [n_out, c_out]
  const int32_t *nbr;    // [K, n_out]
  float *dW;             // [K, c_in, c_out]
  uint32_t c_in, c_out, K, n_out;
  uint32_t G, n_kgroups, mt_cta, n_mtgroups, rows_per_split, n_splits;
  uint32_t blkA, blkB;   // 64-channel blocks per A / B stage
  uint32_t nA, nB, a_stage_bytes, b_stage_bytes;
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) k_wgrad_tc(const WgParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + p.nA * p.a_stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.nA * p.a_stage_bytes +
                                                p.nB * p.b_stage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  auto fullA = [&](uint32_t i) { return bar0 + 8u * i; };
  auto emptyA = [&](uint32_t i) { return bar0 + 8u * (kMaxStages + i); };
  auto fullB = [&](uint32_t i) { return bar0 + 8u * (2 * kMaxStages + i); };
  auto emptyB = [&](uint32_t i) { return bar0 + 8u * (3 * kMaxStages + i); };
  const uint32_t tfull = bar0 + 8u * (4 * kMaxStages);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 * kMaxStages + 4);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item
  uint32_t wi = blockIdx.x;
  const uint32_t kg = wi % p.n_kgroups; wi /= p.n_kgroups;
  const uint32_t mtg = wi % p.n_mtgroups; wi /= p.n_mtgroups;
  const uint32_t split = wi;
  const uint32_t k0 = kg * p.G;
  const uint32_t G = min(p.G, p.K - k0);
  const uint32_t r_begin = split * p.rows_per_split;
  const uint32_t r_end = min(r_begin + p.rows_per_split, p.n_out);
  const uint32_t n_chunks = (r_end - r_begin + kWgRows - 1) / kWgRows;
  const uint32_t ch0 = mtg * p.mt_cta * 128;                 // first input channel of this CTA
  const uint32_t mt = min(p.mt_cta, (p.c_in - ch0 + 127) / 128);  // m-tiles actually present

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < p.nA; ++i) { mbar_init(fullA(i), kProducerWarps * 32); mbar_init(emptyA(i), 1); }
    for (uint32_t i = 0; i < p.nB; ++i) { mbar_init(fullB(i), kProducerWarps * 32); mbar_init(emptyB(i), 1); }
    mbar_init(tfull, 1);
    mbar_fence_init();
  }
  if (warp == kProducerWarps + kEpilogueWarps) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kProducerWarps) {
    const uint32_t sub_row = lane >> 3, j8 = lane & 7;
    const T *in = reinterpret_cast<const T *>(p.in);
    const T *gout = reinterpret_cast<const T *>(p.gout);
    uint32_t g = 0, signaled = 0;
    auto signal = [&](uint32_t gd) {
      fence_proxy_async();
      mbar_arrive(fullA(gd % p.nA));
      if (gd % G == 0) mbar_arrive(fullB((gd / G) % p.nB));
    };
    auto wait_empty = [&](uint32_t bar, uint32_t parity) {
      if (mbar_try_wait(bar, parity)) return;
      cp_async_wait<0>();
      while (signaled < g) signal(signaled++);
      mbar_wait(bar, parity);
    };
    for (uint32_t c = 0; c < n_chunks; ++c) {
      const uint32_t row0 = r_begin + c * kWgRows;
      for (uint32_t kk = 0; kk < G; ++kk) {
        const uint32_t sa = g % p.nA;
        wait_empty(emptyA(sa), ((g / p.nA) & 1u) ^ 1u);
        if (kk == 0) {
          const uint32_t gb = g / G, sb = gb % p.nB;
          wait_empty(emptyB(sb), ((gb / p.nB) & 1u) ^ 1u);
          const uint32_t dstB = sB + sb * p.b_stage_bytes;
          for (uint32_t blk = 0; blk < p.blkB; ++blk) {
            const uint32_t chn = blk * 64 + j8 * 8;
#pragma unroll
            for (int it = 0; it < kWgRows / 16; ++it) {
              const uint32_t row = it * 16 + warp * 4 + sub_row;
              const uint32_t o = row0 + row;
              const bool ok = (o < r_end) && (chn < p.c_out);
              const T *src = gout + (ok ? ((size_t)o * p.c_out + chn) : 0);
              cp_async16(dstB + blk * kBlkBytes + row * 128 + ((j8 ^ (row & 7u)) << 4), src,
                         ok ? 16u : 0u);
            }
          }
        }
        const int32_t *nbr_k = p.nbr + (size_t)(k0 + kk) * p.n_out;
        const uint32_t dstA = sA + sa * p.a_stage_bytes;
        int32_t src_rows[kWgRows / 16];
#pragma unroll
        for (int it = 0; it < kWgRows / 16; ++it) {
          const uint32_t o = row0 + it * 16 + warp * 4 + sub_row;
          src_rows[it] = (o < r_end) ? __ldg(nbr_k + o) : -1;
        }
        for (uint32_t blk = 0; blk < p.blkA; ++blk) {
          const uint32_t chn = ch0 + blk * 64 + j8 * 8;
#pragma unroll
          for (int it = 0; it < kWgRows / 16; ++it) {
            const uint32_t row = it * 16 + warp * 4 + sub_row;
            const bool ok = (src_rows[it] >= 0) && (chn < p.c_in);
            const T *src = in + (ok ? ((size_t)src_rows[it] * p.c_in + chn) : 0);
            cp_async16(dstA + blk * kBlkBytes + row * 128 + ((j8 ^ (row & 7u)) << 4), src,
                       ok ? 16u : 0u);
          }
        }
        cp_async_commit();
        ++g;
        if (g - signaled > (uint32_t)kLag) {
          cp_async_wait<kLag>();
          while (g - signaled > (uint32_t)kLag) signal(signaled++);
        }
      }
    }
    cp_async_wait<0>();
    while (signaled < g) signal(signaled++);
  } else if (warp < kProducerWarps + kEpilogueWarps) {
    const uint32_t q = warp & 3u;
    mbar_wait(tfull, 0);
    tc_fence_after();
    for (uint32_t kk = 0; kk < G; ++kk) {
      for (uint32_t t = 0; t < mt; ++t) {
        const uint32_t ci = ch0 + t * 128 + q * 32 + lane;
        const uint32_t col0 = (kk * p.mt_cta + t) * p.c_out;
        float *dst_row = p.dW + ((size_t)(k0 + kk) * p.c_in + ci) * p.c_out;
        for (uint32_t c = 0; c < p.c_out; c += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((q * 32u) << 16) + col0 + c, v);
          tmem_ld_wait();
          if (ci < p.c_in && n_chunks > 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float f = __uint_as_float(v[i]);
              if (f != 0.f) atomicAdd(dst_row + c + i, f);
            }
          }
        }
      }
    }
  } else if (lane == 0) {
    const uint32_t idesc = umma_idesc_f16(Fmt<T>::v, kTileM, p.c_out, 1, 1);
    const uint32_t lboA = kBlkBytes, lboB = kBlkBytes;
    uint32_t g = 0;
    for (uint32_t c = 0; c < n_chunks; ++c) {
      const uint32_t gb = g / G, sb = gb % p.nB;
      mbar_wait(fullB(sb), (gb / p.nB) & 1u);
      const uint32_t bbase = sB + sb * p.b_stage_bytes;
      for (uint32_t kk = 0; kk < G; ++kk, ++g) {
        const uint32_t sa = g % p.nA;
        mbar_wait(fullA(sa), (g / p.nA) & 1u);
        tc_fence_after();
        const uint32_t abase = sA + sa * p.a_stage_bytes;
        for (uint32_t t = 0; t < mt; ++t) {
          const uint32_t d = tmem_base + (kk * p.mt_cta + t) * p.c_out;
#pragma unroll
          for (int j = 0; j < kWgRows / 16; ++j) {
            uint64_t da = umma_desc(abase + t * 2 * kBlkBytes + j * 2048, lboA, 1024, 2);
            uint64_t db = umma_desc(bbase + j * 2048, lboB, 1024, 2);
            umma_f16(d, da, db, idesc, (c | (uint32_t)j) != 0u);
          }
        }
        umma_commit(emptyA(sa));
      }
      umma_commit(emptyB(sb));
    }
    umma_commit(tfull);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarps + kEpilogueWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

bool conv_wgrad_tc_supported(int dtype, uint32_t c_in, uint32_t c_out) {
  if (dtype != MEB200_BF16 && dtype != MEB200_F16) return false;
  return c_in % 8 == 0 && c_in >= 16 && c_out % 16 == 0 && c_out >= 16 && c_out <= 256;
}

int conv_wgrad_tc(const void *in, const void *grad_out, int dtype, uint32_t c_in, uint32_t K,
                  uint32_t c_out, const int32_t *out_nbr, uint32_t n_out, float *grad_weight,
                  cudaStream_t stream) {
  MEB_CUDA(cudaMemsetAsync(grad_weight, 0, (size_t)K * c_in * c_out * sizeof(float), stream));
  if (n_out == 0 || K == 0) return MEB200_OK;
  WgParams p{};
  p.in = in; p.gout = grad_out; p.nbr = out_nbr; p.dW = grad_weight;
  p.c_in = c_in; p.c_out = c_out; p.K = K; p.n_out = n_out;
  const uint32_t mt_total = cdiv(c_in, 128);
  p.mt_cta = mt_total < kTmemCols / c_out ? mt_total : kTmemCols / c_out;
  if (p.mt_cta > 2) p.mt_cta = 2;
  p.n_mtgroups = cdiv(mt_total, p.mt_cta);
  p.G = kTmemCols / (p.mt_cta * c_out);
  if (p.G > K) p.G = K;
  if (p.G > 8) p.G = 8;
  p.n_kgroups = cdiv(K, p.G);
  p.blkA = p.mt_cta * 2;
  p.blkB = cdiv(c_out, 64);
  p.a_stage_bytes = p.blkA * kBlkBytes;
  p.b_stage_bytes = p.blkB * kBlkBytes;
  const uint32_t bar_bytes = (4 * kMaxStages + 4) * 8 + 16;
  p.nB = 2;
  uint32_t avail = kSmemBudget - p.nB * p.b_stage_bytes - bar_bytes - 1024;
  p.nA = avail / p.a_stage_bytes;
  if (p.nA > (uint32_t)kMaxStages) p.nA = kMaxStages;
  MEB_CHECK_ARG(p.nA >= 4, "not enough shared memory for the wgrad A pipeline");
  // row slices: ~2 waves of CTAs, each slice a multiple of the stage height
  uint32_t base = p.n_kgroups * p.n_mtgroups;
  uint32_t want = cdiv(2ull * num_sms(), base);
  uint32_t max_splits = cdiv(n_out, 4 * kWgRows);
  uint32_t splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
  p.rows_per_split = cdiv(cdiv(n_out, splits), kWgRows) * kWgRows;
  p.n_splits = cdiv(n_out, p.rows_per_split);
  size_t smem_bytes = 1024 + (size_t)p.nA * p.a_stage_bytes + (size_t)p.nB * p.b_stage_bytes + bar_bytes;
  unsigned grid = base * p.n_splits;
  if (dtype == MEB200_BF16) {
    static bool cfg = false;
    if (!cfg) { MEB_CUDA(cudaFuncSetAttribute(k_wgrad_tc<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); cfg = true; }
    k_wgrad_tc<__nv_bfloat16><<<grid, kThreads, smem_bytes, stream>>>(p);
  } else {
    static bool cfg = false;
    if (!cfg) { MEB_CUDA(cudaFuncSetAttribute(k_wgrad_tc<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); cfg = true; }
    k_wgrad_tc<__half><<<grid, kThreads, smem_bytes, stream>>>(p);
  }
  count_tc_launch();
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

}  // namespace meb200
