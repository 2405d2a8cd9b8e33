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
#include <cuda.h>
#include <stdlib.h>
#include <type_traits>
#include <string.h>

#include "conv_tc.cuh"
#include "ptx.cuh"
#include "tc_config.h"

namespace meb200 {

using namespace ptx;

using tc::kTileM;
using tc::kMaxStages;
using tc::kTmemCols;
using tc::kWgRows;
using tc::kBlkBytes;
constexpr int kProducerWarps = 4;   // wgrad kernel
constexpr int kEpilogueWarps = 4;
constexpr int kThreads = (kProducerWarps + kEpilogueWarps + 1) * 32;

struct TcParams {
  const void *A;        // [n_a, c_red] gathered operand
  const void *Wb;       // [K, c_cols, c_red] operand B, reduction dim contiguous
  const int32_t *nbr;   // [K, n_rows]
  void *out;            // [n_rows, c_cols]
  uint32_t c_red, c_cols, K, n_rows;
  uint32_t n_a;          // rows of A (gather indices >= n_a are out of bounds -> zero rows)
  uint32_t wb_row0;      // first row of this launch's column slice inside one offset of Wb
  uint32_t wb_rows_per_k;  // rows of Wb per offset (= full c_cols of the layer)
  uint32_t n_super, R, acc_stages, nA, nB;
  uint32_t a_stage_bytes, b_stage_bytes;
  uint32_t cps, n_macro, a_sub_bytes, b_sub_bytes;  // chunks per stage, stages per offset
  uint32_t out_f32;
  uint32_t lag;          // cp.async groups each producer thread keeps in flight
  uint32_t wb_k_stride;  // elements between consecutive offsets in Wb
  uint32_t out_ld;       // leading dimension (elements) of `out`
};

template <int BK>
struct Swz {
  static constexpr int kRowBytes = BK * 2;
  static constexpr int kChunks = kRowBytes / 16;          // 16-byte chunks per row
  static constexpr int kRowsPerInstr = 32 / kChunks;      // rows one warp-wide cp.async covers
  static constexpr int kItersA = kTileM / kRowsPerInstr;  // warp-wide copies per 128-row tile
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

// barrier slots: fullA[kMax] emptyA[kMax] fullB[kMax] emptyB[kMax] tfull[2] tempty[2]
__device__ __forceinline__ uint32_t bar_fullA(uint32_t b0, uint32_t i) { return b0 + 8u * i; }
__device__ __forceinline__ uint32_t bar_emptyA(uint32_t b0, uint32_t i) { return b0 + 8u * (kMaxStages + i); }
__device__ __forceinline__ uint32_t bar_fullB(uint32_t b0, uint32_t i) { return b0 + 8u * (2 * kMaxStages + i); }
__device__ __forceinline__ uint32_t bar_emptyB(uint32_t b0, uint32_t i) { return b0 + 8u * (3 * kMaxStages + i); }
__device__ __forceinline__ uint32_t bar_tfull(uint32_t b0, uint32_t i) { return b0 + 8u * (4 * kMaxStages + i); }
__device__ __forceinline__ uint32_t bar_tempty(uint32_t b0, uint32_t i) { return b0 + 8u * (4 * kMaxStages + 2 + i); }

// Epilogue warps: accumulator (TMEM) -> registers -> converted 16-byte global stores.
template <typename T, int RT>
__device__ __forceinline__ void epilogue_role(const TcParams &p, uint32_t bar0, uint32_t tmem_base,
                                              uint32_t q, uint32_t lane) {
  constexpr uint32_t R = RT;
  uint32_t t = 0;
  for (uint32_t st = blockIdx.x; st < p.n_super; st += gridDim.x, ++t) {
    const uint32_t acc = t & 1u, par = (t >> 1) & 1u;   // two accumulator sets
    mbar_wait_park(bar_tfull(bar0, acc), par);
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
                                                     (size_t)row * p.out_ld + c);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          } else {
            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<T *>(p.out) +
                                                   (size_t)row * p.out_ld + c);
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
    mbar_arrive(bar_tempty(bar0, acc));
  }
}

// MMA issuer: the WHOLE warp runs the loop with warp-uniform state (ring
// addresses advance by adds, descriptor words are 32-bit adds of compile-time offsets) and one
// elected lane issues.  A single thread can retire only ~1 dependent instruction per 4-6 cycles,
// so every instruction here is on the critical path of the tensor pipe (ncu: the first version,
// ~250 instructions per stage, held the pipe to ~13% busy).
template <typename T, int BK, int RT, int CPS>
__device__ __forceinline__ void mma_role_fast(const TcParams &p, uint32_t sA, uint32_t sB,
                                              uint32_t bar0, uint32_t tmem_base) {
  using S = Swz<BK>;
  constexpr uint32_t R = RT;
  constexpr uint32_t kHi = umma_desc_hi(S::kSbo, S::kLayout);
  constexpr uint32_t kASub16 = (kTileM * BK * 2) >> 4;
  const uint32_t idesc = umma_idesc_f16(Fmt<T>::v, kTileM, p.c_cols, 0, 0);
  const uint32_t a_lo0 = umma_desc_lo(sA, 16), b_lo0 = umma_desc_lo(sB, 16);
  const uint32_t a_inc = p.a_stage_bytes >> 4, b_inc = p.b_stage_bytes >> 4, b_sub16 = p.b_sub_bytes >> 4;
  uint32_t a_lo = a_lo0, b_lo = b_lo0;
  uint32_t sa = 0, pha = 0, sb = 0, phb = 0, t = 0;
  const uint32_t n_groups = p.K * p.n_macro;
  for (uint32_t st = blockIdx.x; st < p.n_super; st += gridDim.x, ++t) {
    const uint32_t acc = t & 1u, par = (t >> 1) & 1u;
    mbar_wait(bar_tempty(bar0, acc), par ^ 1u);
    tc_fence_after();
    const uint32_t d0 = tmem_base + acc * R * p.c_cols;
    for (uint32_t grp = 0; grp < n_groups; ++grp) {
      mbar_wait(bar_fullB(bar0, sb), phb);
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        mbar_wait(bar_fullA(bar0, sa), pha);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = d0 + r * p.c_cols;
#pragma unroll
          for (int sc = 0; sc < CPS; ++sc) {
#pragma unroll
            for (int j = 0; j < BK / 16; ++j) {
              const uint64_t da = pack_desc(a_lo + sc * kASub16 + j * 2, kHi);
              const uint64_t db = pack_desc(b_lo + sc * b_sub16 + j * 2, kHi);
              umma_f16(d, da, db, idesc, (sc | j) != 0 ? 1u : (grp != 0u));
            }
          }
          umma_commit(bar_emptyA(bar0, sa));
          if (r == RT - 1) umma_commit(bar_emptyB(bar0, sb));
        }
        __syncwarp();
        a_lo += a_inc;
        if (++sa == p.nA) { sa = 0; pha ^= 1u; a_lo = a_lo0; }
      }
      b_lo += b_inc;
      if (++sb == p.nB) { sb = 0; phb ^= 1u; b_lo = b_lo0; }
    }
    if (elect_one()) umma_commit(bar_tfull(bar0, acc));
    __syncwarp();
  }
}

// Flow-control words live in shared memory; access them with shared-space volatile ld/st (a
// generic `volatile` pointer compiles to LD.E.STRONG.SYS, a far more expensive access).
__device__ __forceinline__ uint32_t lds_volatile(const volatile uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void *)p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile(volatile uint32_t *p, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32((const void *)p)), "r"(v) : "memory");
}

// Flow control for the register-staged kernels.  The producer warps do not wait on the
// "empty" mbarriers themselves: with more producer warps than ring slots a one-bit parity wait
// could be satisfied by a phase two uses back.  Instead ONE thread (lane 1 of the MMA warp)
// consumes the empty-barrier completions strictly in order and publishes monotonically
// increasing counters in shared memory; a producer may fill stage g once consumed + ring > g.
template <int RT>
__device__ __forceinline__ void release_role(const TcParams &p, uint32_t bar0,
                                             volatile uint32_t *consumed /* [0]=A stages, [1]=B */) {
  uint32_t sa = 0, pha = 0, sb = 0, phb = 0, nA_done = 0, nB_done = 0;
  for (uint32_t st = blockIdx.x; st < p.n_super; st += gridDim.x) {
    for (uint32_t k = 0; k < p.K; ++k) {
      for (uint32_t mc = 0; mc < p.n_macro; ++mc) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          mbar_wait(bar_emptyA(bar0, sa), pha);
          __threadfence_block();
          sts_volatile(consumed, ++nA_done);
          if (++sa == p.nA) { sa = 0; pha ^= 1u; }
        }
        mbar_wait(bar_emptyB(bar0, sb), phb);
        __threadfence_block();
        sts_volatile(consumed + 1, ++nB_done);
        if (++sb == p.nB) { sb = 0; phb ^= 1u; }
      }
    }
  }
}

// =====================================================================================
// Register-staged gather kernel (default forward/dgrad path).
//
// Measured on B200 (profiles/r1_gather_bw_microbench.txt): random row gathers are bounded by
// latency x bytes in flight (~1800 cycles from L2, ~4400 from HBM under load); the TMA
// row-gather sustains only ~6-8 B/cycle/SM and cp.async keeps zero rows in the ring.  Here the
// bytes in flight live in REGISTERS: 12 producer warps each own a pipeline stage, compact the
// valid (tile row, source row) pairs of that stage with ballots, issue up to 16 x 16-byte
// loads per lane for VALID rows only, then write the tile (zero-filled first) into the
// swizzled stage with st.shared and publish it (fence.proxy.async + mbarrier arrive).  The B
// slice comes through the TMA unit (one tile load per channel chunk).
// =====================================================================================
constexpr int kRsProducerWarps = 10;
constexpr int kRsThreads = (kRsProducerWarps + kEpilogueWarps + 2) * 32;   // + MMA warp + release warp
constexpr int kRsBatch = 16;   // 16-byte loads in flight per lane (18 with 3 chunks per stage)

// Wait until `*ctr + ring > g` (slot of stage g recycled): lane 0 polls with back-off, the rest
// of the warp parks at the __syncwarp instead of hammering shared memory.
__device__ __forceinline__ void wait_consumed(volatile uint32_t *ctr, uint32_t g, uint32_t ring,
                                              uint32_t lane) {
  if (lane == 0) {
    while ((int32_t)(g - lds_volatile(ctr)) >= (int32_t)ring) __nanosleep(20);
  }
  __syncwarp();
  __threadfence_block();
}

__device__ __forceinline__ uint4 ldg128(const void *p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T, int BK, int RT, int CPS>
__global__ void __launch_bounds__(kRsThreads, 1)
k_conv_rs(const __grid_constant__ CUtensorMap mapB, const TcParams p) {
  using S = Swz<BK>;
  constexpr int NP = kRsProducerWarps;
  constexpr uint32_t R = RT;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sA = smem_u32(smem);
  const uint32_t sB = sA + p.nA * p.a_stage_bytes;
  uint8_t *tail = smem + p.nA * p.a_stage_bytes + p.nB * p.b_stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
  const uint32_t bar0 = smem_u32(bars);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 * kMaxStages + 4);
  // per-warp compaction scratch: 128 x (source row, tile row); then the flow-control counters
  uint2 *scratch_all = reinterpret_cast<uint2 *>(tail + tc::kBarBytes);
  volatile uint32_t *consumed = reinterpret_cast<volatile uint32_t *>(
      tail + tc::kBarBytes + kRsProducerWarps * kTileM * 8);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < p.nA; ++i) { mbar_init(bar_fullA(bar0, i), 1); mbar_init(bar_emptyA(bar0, i), 1); }
    for (uint32_t i = 0; i < p.nB; ++i) { mbar_init(bar_fullB(bar0, i), 1); mbar_init(bar_emptyB(bar0, i), 1); }
    for (uint32_t i = 0; i < 2; ++i) { mbar_init(bar_tfull(bar0, i), 1); mbar_init(bar_tempty(bar0, i), kEpilogueWarps * 32); }
    mbar_fence_init();
    consumed[0] = 0;
    consumed[1] = 0;
  }
  if (warp == NP + kEpilogueWarps) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < NP) {
    const uint32_t sub_row = lane / S::kChunks, chunk = lane % S::kChunks;
    const T *A = reinterpret_cast<const T *>(p.A);
    uint2 *scratch = scratch_all + warp * kTileM;
    const uint32_t scratch_s = smem_u32(scratch);
    const uint32_t SPK = p.n_macro * R;
    const uint32_t SPS = p.K * SPK;
    const uint32_t n_my = p.n_super > blockIdx.x
                              ? (p.n_super - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t total = n_my * SPS;
    struct Pos { uint32_t st, k, mc, r; };
    auto decode = [&](uint32_t g) {
      Pos q;
      const uint32_t t = g / SPS, rem = g - t * SPS;
      q.st = blockIdx.x + t * gridDim.x;
      q.k = rem / SPK;
      const uint32_t rem2 = rem - q.k * SPK;
      q.mc = rem2 / R;
      q.r = rem2 - q.mc * R;
      return q;
    };
    int32_t idx_cur[4], idx_nxt[4];
    auto load_idx = [&](uint32_t g, int32_t (&dst)[4]) {
      const Pos q = decode(g);
      const int32_t *nbr_k = p.nbr + (size_t)q.k * p.n_rows;
      const uint32_t row0 = (q.st * R + q.r) * kTileM + lane;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t grow = row0 + i * 32;
        dst[i] = (grow < p.n_rows) ? __ldg(nbr_k + grow) : -1;
      }
    };
    constexpr uint32_t step = NP;   // every producer warp takes stages (see release_role)
    uint32_t g = warp;
    if (warp == 0 && lane == 0) tma_prefetch_desc(&mapB);
    if (g < total) load_idx(g, idx_cur);
    const uint32_t lane_lt = (1u << lane) - 1u;
    for (; g < total; g += step) {
      if (g + step < total) load_idx(g + step, idx_nxt);
      const Pos q = decode(g);
      const uint32_t sa = g % p.nA;
      const uint32_t c0 = q.mc * p.cps * BK;
      // ---- compact the valid rows of this stage into the warp's scratch list ----------
      uint32_t V = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ok = idx_cur[i] >= 0;
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (ok) scratch[V + __popc(m & lane_lt)] = make_uint2((uint32_t)idx_cur[i], (uint32_t)(i * 32 + lane));
        V += __popc(m);
      }
      __syncwarp();
      const uint32_t dstA = sA + sa * p.a_stage_bytes;
      bool slot_ready = false;
      constexpr int kRowIters = (CPS == 3 ? 18 : kRsBatch) / CPS;      // row groups per batch
      constexpr uint32_t kRowsPerBatch = kRowIters * S::kRowsPerInstr;
      const T *srcA = A + c0 + chunk * 8;
      uint32_t e0 = 0;
      do {
        // loads for VALID rows only go out first (registers are the landing zone) ...
        uint4 v[kRowIters * CPS];
#pragma unroll
        for (int ri = 0; ri < kRowIters; ++ri) {
          const uint32_t e = e0 + ri * S::kRowsPerInstr + sub_row;
          if (e < V) {
            uint32_t srow;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(srow) : "r"(scratch_s + e * 8));
            const T *src = srcA + (size_t)srow * p.c_red;
#pragma unroll
            for (int sc = 0; sc < CPS; ++sc) {
              v[ri * CPS + sc] = ldg128(src + sc * BK);
            }
          } else {
#pragma unroll
            for (int sc = 0; sc < CPS; ++sc) v[ri * CPS + sc] = make_uint4(0, 0, 0, 0);
          }
        }
        // ... and only then take the ring slot: wait for it, zero it, fill it
        if (!slot_ready) {
          wait_consumed(consumed, g, p.nA, lane);   // ring slot still in use?
          if (q.r == 0 && lane == 0) {
            const uint32_t gb = g / R, sb = gb % p.nB;
            while ((int32_t)(gb - lds_volatile(consumed + 1)) >= (int32_t)p.nB) __nanosleep(20);
            __threadfence_block();
            mbar_arrive_expect_tx(bar_fullB(bar0, sb), CPS * p.c_cols * BK * 2);
            const uint32_t dstB = sB + sb * p.b_stage_bytes;
            const int32_t brow = (int32_t)(q.k * p.wb_rows_per_k + p.wb_row0);
            for (uint32_t sbi = 0; sbi < (uint32_t)CPS; ++sbi)
              tma_load_2d(dstB + sbi * p.b_sub_bytes, &mapB, (int32_t)(c0 + sbi * BK), brow,
                          bar_fullB(bar0, sb));
          }
          for (uint32_t off = lane * 16; off < p.a_stage_bytes; off += 512)
            sts128(dstA + off, make_uint4(0, 0, 0, 0));
          __syncwarp();
          slot_ready = true;
        }
#pragma unroll
        for (int ri = 0; ri < kRowIters; ++ri) {
          const uint32_t e = e0 + ri * S::kRowsPerInstr + sub_row;
          if (e < V) {
            uint32_t trow;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(trow) : "r"(scratch_s + e * 8 + 4));
            const uint32_t dst = dstA + S::offset(trow, chunk);
#pragma unroll
            for (int sc = 0; sc < CPS; ++sc) sts128(dst + sc * p.a_sub_bytes, v[ri * CPS + sc]);
          }
        }
        e0 += kRowsPerBatch;
      } while (e0 < V);
      fence_proxy_async();          // this lane's st.shared -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_fullA(bar0, sa));
#pragma unroll
      for (int i = 0; i < 4; ++i) idx_cur[i] = idx_nxt[i];
    }
  } else if (warp < NP + kEpilogueWarps) {
    epilogue_role<T, RT>(p, bar0, tmem_base, warp & 3u, lane);
  } else if (warp == NP + kEpilogueWarps) {
    mma_role_fast<T, BK, RT, CPS>(p, sA, sB, bar0, tmem_base);   // whole warp, one lane issues
  } else if (lane == 0) {
    // its own warp: a thread parked in mbarrier.try_wait must not hold up the MMA issuer
    release_role<RT>(p, bar0, consumed);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NP + kEpilogueWarps) {
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

// Per-optimizer-step weight packing: fp32 master W[k][ci][co] -> feature dtype in the four
// operand-B layouts of the tensor-core kernels, in one pass (replaces a torch cast plus
// k_transpose_w on EVERY forward call):
//   Wc  [k][ci][co]          dgrad, k_conv_rs          Wt  [k][co][ci]          forward, k_conv_rs
//   Wcp [k][ci][perm(co)]    dgrad, k_conv_ta          Wtp [k][co][perm(ci)]    forward, k_conv_ta
// perm = the order in which k_conv_ta's producers lay a 32-channel block out in tensor memory.
__host__ __device__ constexpr uint32_t ta_perm(uint32_t pos) {   // position -> channel, within 32
  return 8u * ((pos >> 2) & 3u) + 4u * ((pos >> 4) & 1u) + (pos & 3u);
}
template <typename T>
__global__ void __launch_bounds__(256)
k_pack_w(const float *__restrict__ W, T *__restrict__ Wc, T *__restrict__ Wt, T *__restrict__ Wcp,
         T *__restrict__ Wtp, uint32_t c_in, uint32_t c_out) {
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * c_in * c_out;
  const uint32_t ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (uint32_t i = ty; i < 32; i += 8)
    if (ci0 + i < c_in && co0 + tx < c_out) {
      const float v = W[base + (size_t)(ci0 + i) * c_out + co0 + tx];
      tile[i][tx] = v;
      Wc[base + (size_t)(ci0 + i) * c_out + co0 + tx] = from_f32<T>(v);
    }
  __syncthreads();
  const uint32_t px = ta_perm(tx);
  for (uint32_t i = ty; i < 32; i += 8) {
    if (co0 + i < c_out && ci0 + tx < c_in)
      Wt[base + (size_t)(co0 + i) * c_in + ci0 + tx] = from_f32<T>(tile[tx][i]);
    if (Wtp != nullptr && co0 + i < c_out)      // c_in % 32 == 0: the tile is a full block
      Wtp[base + (size_t)(co0 + i) * c_in + ci0 + tx] = from_f32<T>(tile[px][i]);
    if (Wcp != nullptr && ci0 + i < c_in)       // c_out % 32 == 0
      Wcp[base + (size_t)(ci0 + i) * c_out + co0 + tx] = from_f32<T>(tile[i][px]);
  }
}

// All layers of a network in ONE launch (the weights of every layer change together, at the
// optimizer step): job j re-packs one [K, c_in, c_out] tensor exactly as k_pack_w does; a CTA
// finds its job by binary search over the jobs' first-tile indices.
template <typename T>
__global__ void __launch_bounds__(256)
k_pack_w_batched(const PackJob *__restrict__ jobs, uint32_t n_jobs) {
  __shared__ float tile[32][33];
  uint32_t lo = 0, hi = n_jobs - 1;
  while (lo < hi) {                       // last job whose tile_begin <= blockIdx.x
    const uint32_t mid = (lo + hi + 1) >> 1;
    if (jobs[mid].tile_begin <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackJob job = jobs[lo];
  const uint32_t c_in = job.c_in, c_out = job.c_out;
  const uint32_t tx_n = (c_out + 31u) / 32u, ty_n = (c_in + 31u) / 32u;
  uint32_t t = blockIdx.x - job.tile_begin;
  const uint32_t bx = t % tx_n; t /= tx_n;
  const uint32_t by = t % ty_n, bz = t / ty_n;
  const float *W = reinterpret_cast<const float *>(job.w);
  T *Wc = reinterpret_cast<T *>(job.w_cast), *Wt = reinterpret_cast<T *>(job.w_t);
  T *Wcp = reinterpret_cast<T *>(job.w_cp), *Wtp = reinterpret_cast<T *>(job.w_tp);
  const size_t base = (size_t)bz * c_in * c_out;
  const uint32_t ci0 = by * 32, co0 = bx * 32;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (uint32_t i = ty; i < 32; i += 8)
    if (ci0 + i < c_in && co0 + tx < c_out) {
      const float v = W[base + (size_t)(ci0 + i) * c_out + co0 + tx];
      tile[i][tx] = v;
      Wc[base + (size_t)(ci0 + i) * c_out + co0 + tx] = from_f32<T>(v);
    }
  __syncthreads();
  const uint32_t px = ta_perm(tx);
  for (uint32_t i = ty; i < 32; i += 8) {
    if (co0 + i < c_out && ci0 + tx < c_in)
      Wt[base + (size_t)(co0 + i) * c_in + ci0 + tx] = from_f32<T>(tile[tx][i]);
    if (Wtp != nullptr && co0 + i < c_out)
      Wtp[base + (size_t)(co0 + i) * c_in + ci0 + tx] = from_f32<T>(tile[px][i]);
    if (Wcp != nullptr && ci0 + i < c_in)
      Wcp[base + (size_t)(ci0 + i) * c_out + co0 + tx] = from_f32<T>(tile[i][px]);
  }
}

int conv_pack_weights_batched(const PackJob *jobs_dev, uint32_t n_jobs, uint32_t total_tiles,
                              int dtype, cudaStream_t stream) {
  if (n_jobs == 0 || total_tiles == 0) return MEB200_OK;
  if (dtype == MEB200_BF16)
    k_pack_w_batched<__nv_bfloat16><<<total_tiles, 256, 0, stream>>>(jobs_dev, n_jobs);
  else
    k_pack_w_batched<__half><<<total_tiles, 256, 0, stream>>>(jobs_dev, n_jobs);
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

int conv_pack_weights(const float *W, uint32_t K, uint32_t c_in, uint32_t c_out, int dtype,
                      void *w_cast, void *w_t, void *w_cp, void *w_tp, cudaStream_t stream) {
  if (c_out % 32 != 0) w_cp = nullptr;
  if (c_in % 32 != 0) w_tp = nullptr;
  dim3 grid(cdiv(c_out, 32), cdiv(c_in, 32), K);
  if (dtype == MEB200_BF16)
    k_pack_w<__nv_bfloat16><<<grid, 256, 0, stream>>>(
        W, (__nv_bfloat16 *)w_cast, (__nv_bfloat16 *)w_t, (__nv_bfloat16 *)w_cp,
        (__nv_bfloat16 *)w_tp, c_in, c_out);
  else
    k_pack_w<__half><<<grid, 256, 0, stream>>>(W, (__half *)w_cast, (__half *)w_t, (__half *)w_cp,
                                               (__half *)w_tp, c_in, c_out);
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

static int pick_bk(uint32_t c_red) {
  if (c_red % 64 == 0) return 64;
  if (c_red % 32 == 0) return 32;
  if (c_red % 16 == 0) return 16;
  return 0;
}

// ---- TMA tensor maps (driver entry point fetched through the runtime: no libcuda link) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D row-major [rows, cols] 16-bit tensor, box = box_rows x box_cols, swizzle by box width.
static int make_map_2d(CUtensorMap *m, int dtype, const void *base, uint64_t rows, uint64_t cols,
                       uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return MEB200_ERR_UNSUPPORTED;
  }
  // cuTensorMapEncodeTiled is a DRIVER call: it needs a context current on the calling thread.
  // Runtime calls bind the primary context implicitly, driver calls do not, and autograd's
  // backward threads may reach this point before any runtime call (observed: error 201 from
  // the first dgrad of a process).
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    MEB_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (box_cols * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(m, dtype == MEB200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                          : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                  2, const_cast<void *>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box=%ux%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols);
    return MEB200_ERR_CUDA;
  }
  return MEB200_OK;
}

// Opt a kernel in to >48 KB of dynamic shared memory.  The attribute is per (function, device):
// the flags are keyed on the function pointer VALUE and the current device (round 1 kept one
// static flag inside a generic lambda, which every k_wgrad_tc<T,PF> instantiation shared because
// they all decay to the same pointer type -> the second variant launched unconfigured).
static int ensure_big_smem(const void *fn) {
  constexpr int kMaxDev = 16, kMaxFn = 256;
  static const void *fns[kMaxFn];
  static uint16_t done[kMaxFn];
  static int n_fns = 0;
  int dev = 0;
  MEB_CUDA(cudaGetDevice(&dev));
  int slot = -1;
  for (int i = 0; i < n_fns; ++i) if (fns[i] == fn) { slot = i; break; }
  if (slot >= 0 && dev < kMaxDev && (done[slot] >> dev & 1)) return MEB200_OK;
  MEB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  if (slot < 0 && n_fns < kMaxFn) { slot = n_fns++; fns[slot] = fn; done[slot] = 0; }
  if (slot >= 0 && dev < kMaxDev) done[slot] |= (uint16_t)(1u << dev);
  return MEB200_OK;
}
#define MEB_BIG_SMEM(kern) do { int rc__ = ensure_big_smem((const void *)(kern)); if (rc__ != MEB200_OK) return rc__; } while (0)

template <typename T, int BK, int RT, int CPS>
static int launch_rs(const CUtensorMap &mB, const TcParams &p, size_t smem_bytes, unsigned grid,
                     cudaStream_t stream) {
  auto kern = k_conv_rs<T, BK, RT, CPS>;
  MEB_BIG_SMEM(kern);
  kern<<<grid, kRsThreads, smem_bytes, stream>>>(mB, p);
  count_tc_launch();
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

template <typename T>
static int launch_rs_bk_r(int bk, uint32_t R, uint32_t cps, const CUtensorMap &mB, const TcParams &p,
                          size_t smem, unsigned grid, cudaStream_t s) {
#define MEB_RS_R(BKV, CP)                                                  \
  switch (R) {                                                             \
    case 4: return launch_rs<T, BKV, 4, CP>(mB, p, smem, grid, s);         \
    case 2: return launch_rs<T, BKV, 2, CP>(mB, p, smem, grid, s);         \
    default: return launch_rs<T, BKV, 1, CP>(mB, p, smem, grid, s);        \
  }
  if (bk == 64) {
    if (cps == 2) { MEB_RS_R(64, 2) }
    MEB_RS_R(64, 1)
  }
  if (bk == 32) {
    if (cps == 3) { MEB_RS_R(32, 3) }
    MEB_RS_R(32, 1)
  }
  if (cps == 3) { MEB_RS_R(16, 3) }
  MEB_RS_R(16, 1)
#undef MEB_RS_R
}

// One launch of k_conv_rs over output columns [0, c_cols) of a problem whose B operand / output
// may be slices of wider tensors (N-splitting for c_cols > 256).  The fallback for reduction
// widths that are not a multiple of 32 channels and for callers that pass un-packed weights.
static int launch_tc_slice(const void *A, int dtype, uint32_t n_a, uint32_t c_reduce,
                           const void *Wb_full, uint32_t c_cols_full, uint32_t n0, uint32_t K,
                           uint32_t c_cols, const int32_t *nbr, uint32_t n_rows, void *out,
                           uint32_t out_ld, int out_dtype, cudaStream_t stream) {
  const void *Wb = reinterpret_cast<const uint8_t *>(Wb_full) + (size_t)n0 * c_reduce * 2;
  TcParams p{};
  p.A = A; p.Wb = Wb; p.nbr = nbr; p.out = out;
  p.c_red = c_reduce; p.c_cols = c_cols; p.K = K; p.n_rows = n_rows;
  p.out_f32 = out_dtype == MEB200_F32;
  p.wb_k_stride = c_cols_full * c_reduce; p.out_ld = out_ld;
  p.n_a = n_a; p.wb_row0 = n0; p.wb_rows_per_k = c_cols_full;
  if ((reinterpret_cast<uintptr_t>(A) & 15) != 0 || (reinterpret_cast<uintptr_t>(Wb_full) & 15) != 0 ||
      encode_tiled_fn() == nullptr) {
    set_error("conv tc: operands must be 16-byte aligned and the driver must provide tensor maps");
    return MEB200_ERR_UNSUPPORTED;
  }
  // 24 KB stages: one 128B-swizzle chunk per stage for BK = 64, three 64B-swizzle chunks for
  // BK = 32 (the best of the round-1 sweep)
  tc::FwdCfg cfg = tc::fwd_config(c_reduce, c_cols, n_rows, 24 * 1024, tc::kRsScratchBytes);
  if (!((cfg.bk == 64 && cfg.cps <= 2) || (cfg.bk != 64 && (cfg.cps == 1 || cfg.cps == 3))))
    cfg = tc::fwd_config(c_reduce, c_cols, n_rows, 1, tc::kRsScratchBytes);   // one chunk per stage
  if (cfg.bk == 0 || cfg.pipe.lag < 1) {
    set_error("conv tc: no pipeline configuration fits (c_red=%u c_cols=%u)", c_reduce, c_cols);
    return MEB200_ERR_UNSUPPORTED;
  }
  const uint32_t R = cfg.R;
  p.R = R; p.acc_stages = 2; p.n_super = cfg.n_super;
  p.a_stage_bytes = cfg.a_stage_bytes; p.b_stage_bytes = cfg.b_stage_bytes;
  p.cps = cfg.cps; p.n_macro = c_reduce / (uint32_t)cfg.bk / cfg.cps;
  p.a_sub_bytes = cfg.a_sub_bytes; p.b_sub_bytes = cfg.b_sub_bytes;
  p.lag = cfg.pipe.lag; p.nA = cfg.pipe.nA; p.nB = cfg.pipe.nB;
  const unsigned grid = p.n_super < (uint32_t)num_sms() ? p.n_super : (unsigned)num_sms();
  CUtensorMap mB;
  memset(&mB, 0, sizeof(mB));
  int rc = make_map_2d(&mB, dtype, Wb_full, (uint64_t)K * c_cols_full, c_reduce, (uint32_t)cfg.bk, c_cols);
  if (rc != MEB200_OK) return rc;
  if (dtype == MEB200_BF16)
    return launch_rs_bk_r<__nv_bfloat16>(cfg.bk, R, cfg.cps, mB, p, cfg.smem_bytes, grid, stream);
  return launch_rs_bk_r<__half>(cfg.bk, R, cfg.cps, mB, p, cfg.smem_bytes, grid, stream);
}

// =====================================================================================
// Operand A in TENSOR MEMORY (k_conv_ta): the default forward/dgrad kernel for bf16/fp16 features
// whose reduction width is a multiple of 32 channels.
//
// k_conv_rs stages the gathered rows in shared memory; per (tile, offset) stage that costs a
// 24 KB zero fill, the valid rows, and the tensor core reading A and B back — ~82 KB of shared
// memory traffic against 288 cycles of tensor time on the 96-channel layers (profiles/r1_notes.md).
// Here the gathered rows never touch shared memory:
//   producers : 3 groups x 4 warps.  A group owns every 3rd (tile, offset, channel chunk) stage;
//               its four warps cover the four 32-lane quadrants of tensor memory.  Thread t of a
//               warp handles rows {t/4, t/4+8, t/4+16, t/4+24} of the quadrant and the t%4-th
//               16-byte piece of every 64-byte block of those rows: a warp-wide load touches 8
//               rows x 64 contiguous bytes (the same sector efficiency as a cooperative row copy),
//               missing neighbours stay zero registers, and `tcgen05.st.16x256b.x2` drops the
//               registers straight into the A operand (TMEM lane = tile row).
//   MMA warp  : tcgen05.mma with A from TMEM, B (the packed weights) from shared memory via TMA.
//   epilogue  : unchanged (tcgen05.ld -> convert -> 16-byte stores).
// Because four threads share a row, TMEM column c of a 32-channel block does not hold channels
// (2c, 2c+1): position p = 16h + 4j + 2e + b of the reduction axis holds channel 8j + 4h + 2e + b
// (h: 8-column group, j: piece/thread, e: word, b: half word).  The reduction order is free as
// long as B agrees, so the weights are PACKED in that order once per optimizer step
// (k_pack_w, ta_perm) instead of shuffling registers in the hot loop.
// =====================================================================================
struct TaParams {
  const void *A;        // [n_a, c_red] gathered operand
  const int32_t *nbr;   // [K, n_rows]
  void *out;            // [n_rows, c_cols] (leading dimension out_ld)
  uint32_t c_red, c_cols, K, n_rows;
  uint32_t wb_row0, wb_rows_per_k;   // row of this launch's column slice inside one offset of Wb
  uint32_t n_super, R, acc_sets, nA, nB;
  uint32_t n_macro;                  // channel chunks (stages) per (tile, offset)
  uint32_t b_stage_bytes, b_sub_bytes;
  uint32_t a_col0;                   // first TMEM column of the A slots
  uint32_t out_f32, out_ld;
  uint32_t rs;                       // k_conv_ts: row slots of a producer thread's ring
  uint32_t Kg;                       // offset groups: K, or ceil(K / G) when a stage spans G offsets
};

__device__ __forceinline__ void tmem_st_16x256b_x2(uint32_t taddr, uint32_t r0, uint32_t r1,
                                                   uint32_t r2, uint32_t r3, uint32_t r4,
                                                   uint32_t r5, uint32_t r6, uint32_t r7) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(r4), "r"(r5), "r"(r6), "r"(r7)
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// ---- roles shared by the two A-in-TMEM kernels ------------------------------------------------
template <typename T>
__device__ __forceinline__ void ta_epilogue(const TaParams &p, uint32_t bar0, uint32_t tmem_base,
                                            uint32_t q, uint32_t lane) {
  const uint32_t R = p.R;
  uint32_t acc = 0, par = 0;
  for (uint32_t st = blockIdx.x; st < p.n_super; st += gridDim.x) {
    // a super tile takes tens of microseconds: poll rarely, the producers need the issue slots
    // (the busy wait was 19 % of all issued instructions in the first capture)
    mbar_wait_park(bar_tfull(bar0, acc), par);
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
                                                     (size_t)row * p.out_ld + c);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                   __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          } else {
            uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<T *>(p.out) +
                                                   (size_t)row * p.out_ld + c);
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
    mbar_arrive(bar_tempty(bar0, acc));
    if (++acc == p.acc_sets) { acc = 0; par ^= 1u; }
  }
}

// MMA issuer: the whole warp runs the loop, one elected lane issues (A from TMEM, B from smem)
template <typename T, int NB, int BK>
__device__ __forceinline__ void ta_mma(const TaParams &p, uint32_t sB, uint32_t bar0,
                                       uint32_t tmem_base) {
  using S = Swz<BK>;
  constexpr uint32_t KC = 32u * NB, kACols = 16u * NB;
  constexpr uint32_t kHi = umma_desc_hi(S::kSbo, S::kLayout);
  const uint32_t R = p.R;
  const uint32_t idesc = umma_idesc_f16(Fmt<T>::v, kTileM, p.c_cols, 0, 0);
  const uint32_t b_lo0 = umma_desc_lo(sB, 16);
  const uint32_t b_inc = p.b_stage_bytes >> 4, b_sub16 = p.b_sub_bytes >> 4;
  uint32_t b_lo = b_lo0, sa = 0, pha = 0, sb = 0, phb = 0, acc = 0, par = 0;
  const uint32_t n_groups = p.Kg * p.n_macro;
  for (uint32_t st = blockIdx.x; st < p.n_super; st += gridDim.x) {
    mbar_wait_park(bar_tempty(bar0, acc), par ^ 1u);
    tc_fence_after();
    const uint32_t d0 = tmem_base + acc * R * p.c_cols;
    for (uint32_t grp = 0; grp < n_groups; ++grp) {
      mbar_wait_park(bar_fullB(bar0, sb), phb);
      for (uint32_t r = 0; r < R; ++r) {
        mbar_wait_park(bar_fullA(bar0, sa), pha);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = d0 + r * p.c_cols;
          const uint32_t a = tmem_base + p.a_col0 + sa * kACols;
#pragma unroll
          for (int s = 0; s < (int)(KC / 16); ++s) {
            constexpr int kStepsPerSub = BK / 16;
            const uint64_t db = pack_desc(b_lo + (s / kStepsPerSub) * b_sub16 + (s % kStepsPerSub) * 2, kHi);
            umma_f16_ts(d, a + s * 8, db, idesc, s != 0 ? 1u : (grp != 0u));
          }
          umma_commit(bar_emptyA(bar0, sa));
          if (r == R - 1) umma_commit(bar_emptyB(bar0, sb));
        }
        __syncwarp();
        if (++sa == p.nA) { sa = 0; pha ^= 1u; }
      }
      b_lo += b_inc;
      if (++sb == p.nB) { sb = 0; phb ^= 1u; b_lo = b_lo0; }
    }
    if (elect_one()) umma_commit(bar_tfull(bar0, acc));
    __syncwarp();
    if (++acc == p.acc_sets) { acc = 0; par ^= 1u; }
  }
}

// G > 1: a stage spans G consecutive offsets (narrow layers: c_red = 32 NB / G channels each),
// its B sub-tiles come from G different offsets; offsets past K repeat the last one (their A
// blocks are zero).
template <int NB, int BK, int G = 1>
__device__ __forceinline__ void ta_bload(const CUtensorMap *mapB, const TaParams &p, uint32_t sB,
                                         uint32_t bar0) {
  constexpr uint32_t KC = 32u * NB;
  constexpr int kSubs = (int)(KC / BK), kSubsPerOffset = kSubs / G;
  tma_prefetch_desc(mapB);
  uint32_t sb = 0, phb = 0;
  for (uint32_t st = blockIdx.x; st < p.n_super; st += gridDim.x) {
    for (uint32_t kg = 0; kg < p.Kg; ++kg) {
      for (uint32_t mc = 0; mc < p.n_macro; ++mc) {
        mbar_wait_park(bar_emptyB(bar0, sb), phb ^ 1u);
        mbar_arrive_expect_tx(bar_fullB(bar0, sb), KC * p.c_cols * 2);
        const uint32_t dstB = sB + sb * p.b_stage_bytes;
#pragma unroll
        for (int sub = 0; sub < kSubs; ++sub) {
          uint32_t k = kg, c0 = mc * KC + sub * BK;
          if constexpr (G > 1) {
            k = min(kg * G + sub / kSubsPerOffset, p.K - 1);
            c0 = (sub % kSubsPerOffset) * BK;
          }
          tma_load_2d(dstB + sub * p.b_sub_bytes, mapB, (int32_t)c0,
                      (int32_t)(k * p.wb_rows_per_k + p.wb_row0), bar_fullB(bar0, sb));
        }
        if (++sb == p.nB) { sb = 0; phb ^= 1u; }
      }
    }
  }
}

// position of a producer in the stage sequence g -> (super tile, offset, channel chunk, row tile),
// advanced without divisions
struct TaPos {
  uint32_t st, k, mc, r;
  __device__ __forceinline__ void init(uint32_t g, uint32_t R, uint32_t n_macro, uint32_t K) {
    const uint32_t SPK = n_macro * R, SPS = K * SPK;
    const uint32_t t = g / SPS, rem = g - t * SPS;
    st = blockIdx.x + t * gridDim.x;
    k = rem / SPK;
    const uint32_t rem2 = rem - k * SPK;
    mc = rem2 / R;
    r = rem2 - mc * R;
  }
  __device__ __forceinline__ void advance(uint32_t by, uint32_t R, uint32_t n_macro, uint32_t K) {
    r += by;
    while (r >= R) {
      r -= R;
      if (++mc == n_macro) {
        mc = 0;
        if (++k == K) { k = 0; st += gridDim.x; }
      }
    }
  }
};

// ---- gathered rows wait in SHARED MEMORY, one private ring per thread ---------------------------
// (A first version kept them in registers: 3 stages x ~7.6 KB of useful loads in flight per SM,
// ~1200 cycles per stage, profiles/r2_notes.md.)  A producer thread copies its VALID rows with
// cp.async into a private ring
// of RS row slots (layout [block][slot][lane] x 16 B: conflict free for the copies and for the
// 128-bit reads back), runs up to kTsDepth stages ahead, and only then moves a landed stage
// smem -> registers -> tensor memory.  Two groups of 4 warps alternate stages; missing rows take
// no ring space, so the ring holds ~8 stages of useful bytes per group.
constexpr int kTsGroups = (int)tc::kTsGroupsCfg;
constexpr int kTsProducerWarps = 4 * kTsGroups;
constexpr int kTsThreads = (kTsProducerWarps + kEpilogueWarps + 2) * 32;   // + MMA warp, weights loader
constexpr uint32_t kTsDepth = 8;             // stage descriptors / landing barriers per warp
constexpr int kTsPf = 4;                     // neighbour indices are copied this many of the group's stages ahead

template <typename T, int NB, int BK, int G>
__global__ void __launch_bounds__(kTsThreads, 1)
k_conv_ts(const __grid_constant__ CUtensorMap mapB, const TaParams p) {
  // G = offsets per stage.  G = 1: a stage is NB 32-channel blocks of ONE offset's rows (chunk
  // mc of c_red).  G > 1 (c_red = 32 NB / G): the stage concatenates the rows of G consecutive
  // offsets along the reduction axis — out = sum_k A_k W_k = [A_k0 .. A_k(G-1)] [W_k0; ..] — so
  // the narrow layers get stages as fat as the wide ones instead of G times as many thin ones.
  // G = 0 (STEM, rows of 4 channels = 8 bytes, e.g. the 3-channel input of a network): the
  // reduction axis is (offset, channel) with 8 offsets per 32-channel block, a stage covers
  // GI = 8 NB offsets of one row tile; everything downstream of the producers sees a K = 1
  // convolution over c_red = 4 * (padded K) virtual channels (p.Kg = 1, p.n_macro = chunks).
  constexpr bool STEM = (G == 0);
  constexpr int GI = STEM ? 8 * NB : G;        // offsets whose indices a stage needs
  constexpr uint32_t KC = 32u * NB;
  constexpr uint32_t kACols = 16u * NB;
  constexpr int NBo = STEM ? 1 : NB / (STEM ? 1 : G);   // 32-channel blocks per offset
  constexpr int NP = kTsProducerWarps;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sB = smem_u32(smem);
  const uint32_t ring_bytes = NB * (p.rs + 1) * 512u;                // per producer warp: RS row slots + the zero slot
  uint8_t *rings = smem + p.nB * p.b_stage_bytes;
  uint8_t *tail = rings + NP * ring_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
  const uint32_t bar0 = smem_u32(bars);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 * kMaxStages + 4);
  uint64_t *landed_all = reinterpret_cast<uint64_t *>(tail + tc::kBarBytes + 64);   // [NP][kTsDepth]
  uint32_t *desc_all = reinterpret_cast<uint32_t *>(landed_all + NP * kTsDepth);    // [NP][kTsDepth][32]
  int32_t *idx_all = reinterpret_cast<int32_t *>(desc_all + NP * kTsDepth * 32);     // [NP][kTsPf][GI][32]

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t R = p.R;
  const uint32_t SPS = p.Kg * p.n_macro * R;
  const uint32_t n_my = p.n_super > blockIdx.x
                            ? (p.n_super - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t total = n_my * SPS;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < p.nA; ++i) { mbar_init(bar_fullA(bar0, i), 4); mbar_init(bar_emptyA(bar0, i), 1); }
    for (uint32_t i = 0; i < p.nB; ++i) { mbar_init(bar_fullB(bar0, i), 1); mbar_init(bar_emptyB(bar0, i), 1); }
    for (uint32_t i = 0; i < 2; ++i) { mbar_init(bar_tfull(bar0, i), 1); mbar_init(bar_tempty(bar0, i), kEpilogueWarps * 32); }
    mbar_fence_init();
  }
  if (warp < NP) {
    // each producer warp initialises its own landing barriers (one thread doing all ~100
    // barriers of the CTA was a visible serial prologue on the 30-60 us layers)
    if (lane < kTsDepth) {
      mbar_init(smem_u32(landed_all + warp * kTsDepth + lane), 32);
      mbar_fence_init();
    }
    // slot RS of every block of this warp's ring holds zeros for good: rows without a neighbour
    // point there, so the read-back path has no branches and no register zeroing
    const uint32_t zbase = smem_u32(rings + warp * ring_bytes) + lane * 16 + p.rs * 512u;
    for (int b = 0; b < NB; ++b) sts128(zbase + b * (p.rs + 1) * 512u, make_uint4(0, 0, 0, 0));
  }
  if (warp == NP + kEpilogueWarps) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < NP) {
    // =============================== PRODUCERS ======================================
    const uint32_t grp = warp >> 2, quad = warp & 3u;
    const uint32_t piece = lane & 3u, rsub = lane >> 2;
    const T *A = reinterpret_cast<const T *>(p.A) + piece * 8;
    const uint32_t ring = smem_u32(rings + warp * ring_bytes) + lane * 16;   // + (b*RS + slot)*512
    const uint32_t landed = smem_u32(landed_all + warp * kTsDepth);
    uint32_t *desc = desc_all + warp * kTsDepth * 32 + lane;                 // + d * 32
    const uint32_t RS = p.rs;
    // stages of this group: j = 0, 1, ...  <->  g = grp + kTsGroups * j
    const uint32_t n_j = total > grp ? (total - grp + kTsGroups - 1) / kTsGroups : 0;
    TaPos pis, ppf;                   // position of the next stage to issue / to prefetch indices for
    pis.init(grp, R, p.n_macro, p.Kg);
    ppf = pis;
    // The neighbour table streams from HBM (no reuse): its latency (~2 us under load) must be
    // covered by the prefetch distance.  Registers cannot do it — rotating a register ring makes
    // every iteration wait for the newest load (measured: 12 % of all stall samples on the first
    // use, r2_run4) — so the indices of stage j + kTsPf are copied by cp.async (4 bytes per lane,
    // one coalesced 128-byte row of the table per warp and offset) into a per-warp ring in shared
    // memory, one commit group per issued stage.
    const uint32_t iring = smem_u32(idx_all + warp * kTsPf * GI * 32);
    auto copy_idx = [&](const TaPos &q, uint32_t j) {
      if (j < n_j) {
        const uint32_t grow = (q.st * R + q.r) * kTileM + quad * 32 + lane;
#pragma unroll
        for (int g = 0; g < GI; ++g) {
          const uint32_t k = (STEM ? q.mc : q.k) * GI + g;
          const bool ok = grow < p.n_rows && k < p.K;
          cp_async4(iring + (((j % kTsPf) * GI + g) * 32 + lane) * 4,
                    p.nbr + (ok ? (size_t)k * p.n_rows + grow : 0), ok ? 4u : 0u);
        }
      }
      cp_async_commit();
    };
#pragma unroll 1
    for (int d = 0; d < kTsPf; ++d) {
      copy_idx(ppf, (uint32_t)d);
      if (d + 1 < kTsPf) ppf.advance(kTsGroups, R, p.n_macro, p.Kg);
    }
    uint32_t head = 0, tail_j = 0;           // consumed / issued stage counts of this group
    uint32_t wr = 0, used = 0;               // this lane's ring: next slot, slots in use
    uint32_t slot_head = grp % p.nA, slot_par = 0;    // TMEM slot of the head stage, its use parity
    while (head < n_j) {
      // ---- run ahead: issue the copies of further stages while every lane has ring space ----
      while (tail_j < n_j && tail_j - head < kTsDepth) {
        // indices of stage tail_j: copied kTsPf issued stages ago (or in the prologue)
        cp_async_wait<kTsPf - 1>();
        __syncwarp();
        uint32_t nv = 0, dw = 0;
        const uint32_t grow0 = (pis.st * R + pis.r) * kTileM + quad * 32 + rsub;
        const uint32_t d = tail_j % kTsDepth;
        const uint32_t bstride = (RS + 1) * 512u;
        if constexpr (STEM) {
          // every row takes a slot (with 16 offsets per stage a row without any neighbour is
          // rare); a 16-byte piece = the 4-channel rows of TWO offsets, 8 bytes each
          nv = 4;
          if (!__all_sync(0xffffffffu, used + nv <= RS)) break;
          const T *A4 = reinterpret_cast<const T *>(p.A);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const uint32_t rq = (m >> 1) * 16 + (m & 1) * 8;
            const bool row_ok = grow0 + rq < p.n_rows;
            const uint32_t dst = ring + wr * 512u;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const uint32_t o = b * 8 + piece * 2 + e;
                int32_t v;
                asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v)
                             : "r"(iring + (((tail_j % kTsPf) * GI + o) * 32 + rq + rsub) * 4));
                const bool ok = row_ok && pis.mc * GI + o < p.K && v >= 0;
                cp_async8(dst + b * bstride + e * 8, A4 + (ok ? (size_t)v * 4 : 0), ok ? 8u : 0u);
              }
            }
            dw |= wr << (8 * m);
            if (++wr == RS) wr = 0;
          }
        } else {
        int32_t srow[4][GI];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const uint32_t rq = (m >> 1) * 16 + (m & 1) * 8;
          bool any = false;
#pragma unroll
          for (int g = 0; g < GI; ++g) {
            int32_t v;
            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v)
                         : "r"(iring + (((tail_j % kTsPf) * GI + g) * 32 + rq + rsub) * 4));
            // rows past the end and offsets past K were zero-filled, not -1
            const bool in_range = grow0 + rq < p.n_rows && pis.k * GI + g < p.K;
            srow[m][g] = in_range ? v : -1;
            any |= srow[m][g] >= 0;
          }
          if (!any) srow[m][0] = -2;          // the row takes no ring slot at all
          nv += any;
        }
        if (!__all_sync(0xffffffffu, used + nv <= RS)) break;
        const T *src0 = A + pis.mc * KC;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          uint32_t code = RS;                 // the zero slot
          if (srow[m][0] != -2) {
            code = wr;
            const uint32_t dst = ring + wr * 512u;
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const int32_t sr = srow[m][b / NBo];
              // G > 1: block b = channels 32 (b % NBo) of offset b / NBo; a missing offset of a
              // row that has others is zero-filled by the copy engine
              const T *src = (GI == 1 ? src0 : A) + (sr < 0 ? 0 : (size_t)sr * p.c_red) +
                             (GI == 1 ? b : b % NBo) * 32;
              cp_async16(dst + b * bstride, src, sr < 0 ? 0u : 16u);
            }
            if (++wr == RS) wr = 0;
          }
          dw |= code << (8 * m);
        }
        }
        used += nv;
        desc[d * 32] = dw;
        cp_async_mbar_arrive(landed + d * 8);
        ++tail_j;
        pis.advance(kTsGroups, R, p.n_macro, p.Kg);
        ppf.advance(kTsGroups, R, p.n_macro, p.Kg);
        __syncwarp();                 // every lane has read stage tail_j - 1's ring row: reuse it
        copy_idx(ppf, tail_j + kTsPf - 1);
      }
      // ---- move the oldest landed stage: shared memory -> registers -> tensor memory ----------
      {
        const uint32_t d = head % kTsDepth;
        mbar_wait_park(landed + d * 8, (head / kTsDepth) & 1u);
        const uint32_t dw = desc[d * 32];
        // read the rows back with 16-byte loads (conflict free in the [block][slot][lane] layout;
        // 8-byte loads in tcgen05.st's register order saved the register moves before the store
        // but doubled the shared-memory wavefronts: 0.62 vs 0.58 ms on the largest layer)
        const uint32_t bstride = (RS + 1) * 512u;
        uint4 v[4][NB];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const uint32_t code = (dw >> (8 * m)) & 0xffu;
          const uint32_t src = ring + code * 512u;          // code == RS: the zero slot
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[m][b].x), "=r"(v[m][b].y), "=r"(v[m][b].z), "=r"(v[m][b].w)
                         : "r"(src + b * bstride));
          }
          used -= code != RS;
        }
        // A slot recycled by the tensor core?  nA is a multiple of the number of groups and the
        // groups take stages in turn, so a slot always belongs to the same group: its previous
        // use is this group's own stage and a one-bit parity wait on the "empty" barrier is exact.
        if (lane == 0) mbar_wait_park(bar_emptyA(bar0, slot_head), slot_par ^ 1u);
        __syncwarp();
        tc_fence_after();
        const uint32_t ta = tmem_base + p.a_col0 + slot_head * kACols + ((quad * 32u) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const uint4 lo = v[2 * half][b], hi = v[2 * half + 1][b];
            tmem_st_16x256b_x2(ta + ((uint32_t)(half * 16) << 16) + b * 16,
                               lo.x, lo.y, hi.x, hi.y, lo.z, lo.w, hi.z, hi.w);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_fullA(bar0, slot_head));
        ++head;
        slot_head += kTsGroups;
        if (slot_head >= p.nA) { slot_head -= p.nA; slot_par ^= 1u; }
      }
    }
    cp_async_wait<0>();
  } else if (warp < NP + kEpilogueWarps) {
    ta_epilogue<T>(p, bar0, tmem_base, warp & 3u, lane);
  } else if (warp == NP + kEpilogueWarps) {
    ta_mma<T, NB, BK>(p, sB, bar0, tmem_base);
  } else if (warp == NP + kEpilogueWarps + 1) {
    if (lane == 0) ta_bload<NB, BK, (STEM ? 1 : G)>(&mapB, p, sB, bar0);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NP + kEpilogueWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <typename T, int NB, int BK, int G>
static int launch_ts(const CUtensorMap &mB, const TaParams &p, size_t smem_bytes, unsigned grid,
                     cudaStream_t stream) {
  auto kern = k_conv_ts<T, NB, BK, G>;
  MEB_BIG_SMEM(kern);
  kern<<<grid, kTsThreads, smem_bytes, stream>>>(mB, p);
  count_tc_launch();
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

// MEB200_TC_KERNEL=rs forces the round-1 kernel (A in shared memory) for A/B measurements.
static bool ts_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("MEB200_TC_KERNEL");
    v = (e && e[0] == 'r') ? 0 : 1;
    if (encode_tiled_fn() == nullptr) v = 0;
  }
  return v == 1;
}

// One launch of k_conv_ts over output columns [n0, n0 + c_cols) (c_cols <= 256).
// Wp = PERMUTED packed weights [K, c_cols_full, c_reduce] (reduction axis in ta_perm order).
static int launch_ta_slice(const void *A, int dtype, uint32_t n_a, uint32_t c_reduce,
                           const void *Wp_full, uint32_t c_cols_full, uint32_t n0, uint32_t K,
                           uint32_t c_cols, const int32_t *nbr, uint32_t n_rows, void *out,
                           uint32_t out_ld, int out_dtype, cudaStream_t stream) {
  (void)n_a;
  // MEB200_TA_R / MEB200_TA_ACC: tuning overrides (row tiles per weight stage, accumulator sets)
  static int force_R = -1, force_acc = -1;
  if (force_R < 0) {
    const char *e = getenv("MEB200_TA_R");
    force_R = e ? atoi(e) : 0;
    e = getenv("MEB200_TA_ACC");
    force_acc = e ? atoi(e) : 0;
  }
  tc::TaCfg cfg = tc::ta_config(c_reduce, c_cols, n_rows, force_R, force_acc);
  if (cfg.nb == 0 && (force_R || force_acc)) cfg = tc::ta_config(c_reduce, c_cols, n_rows);
  if (cfg.nb == 0) return MEB200_ERR_UNSUPPORTED;
  TaParams p{};
  p.A = A; p.nbr = nbr; p.out = out;
  p.c_red = c_reduce; p.c_cols = c_cols; p.K = K; p.n_rows = n_rows;
  p.wb_row0 = n0; p.wb_rows_per_k = c_cols_full;
  p.n_super = cfg.n_super; p.R = cfg.R; p.acc_sets = cfg.acc_sets; p.nA = cfg.nA; p.nB = cfg.nB;
  p.n_macro = cfg.n_macro; p.b_stage_bytes = cfg.b_stage_bytes; p.b_sub_bytes = cfg.b_sub_bytes;
  p.a_col0 = cfg.a_col0; p.out_f32 = out_dtype == MEB200_F32; p.out_ld = out_ld;
  p.rs = cfg.rs;
  p.Kg = (K + cfg.g - 1) / cfg.g;
  CUtensorMap mB;
  memset(&mB, 0, sizeof(mB));
  int rc = make_map_2d(&mB, dtype, Wp_full, (uint64_t)K * c_cols_full, c_reduce, (uint32_t)cfg.bk, c_cols);
  if (rc != MEB200_OK) return rc;
  const unsigned grid = p.n_super < (uint32_t)num_sms() ? p.n_super : (unsigned)num_sms();
#define MEB_TA(TT)                                                                        \
  switch (cfg.nb * 10 + cfg.g) {                                                          \
    case 11: return launch_ts<TT, 1, 32, 1>(mB, p, cfg.smem_bytes, grid, stream);         \
    case 21: return launch_ts<TT, 2, 64, 1>(mB, p, cfg.smem_bytes, grid, stream);         \
    case 31: return launch_ts<TT, 3, 32, 1>(mB, p, cfg.smem_bytes, grid, stream);         \
    case 41: return launch_ts<TT, 4, 64, 1>(mB, p, cfg.smem_bytes, grid, stream);         \
    case 42: return launch_ts<TT, 4, 64, 2>(mB, p, cfg.smem_bytes, grid, stream);         \
    case 44: return launch_ts<TT, 4, 32, 4>(mB, p, cfg.smem_bytes, grid, stream);         \
    default: break;                                                                       \
  }
  if (dtype == MEB200_BF16) { MEB_TA(__nv_bfloat16) } else { MEB_TA(__half) }
#undef MEB_TA
  return MEB200_ERR_UNSUPPORTED;
}

// ---- network stem: rows of 4 channels (8 bytes), see k_conv_ts STEM --------------------------
// A4 = input features padded to 4 channels [n_a, 4]; Wv = permuted packed weights of the virtual
// K = 1 layer, [c_cols, 4 * 16 * ceil(K / 16)] (reduction axis contiguous, ta_perm order inside
// each 32-channel block = 8 offsets x 4 channels).
bool conv_stem_tc_supported(int dtype, uint32_t K, uint32_t c_cols) {
  if (dtype != MEB200_BF16 && dtype != MEB200_F16) return false;
  return ts_enabled() && tc::ta_stem_config(K, c_cols, 1u << 20).nb != 0;
}

int conv_stem_forward_tc(const void *A4, int dtype, uint32_t K, const void *Wv, uint32_t c_cols,
                         const int32_t *nbr, uint32_t n_rows, void *out, int out_dtype,
                         cudaStream_t stream) {
  if (n_rows == 0) return MEB200_OK;
  const tc::TaCfg cfg = tc::ta_stem_config(K, c_cols, n_rows);
  if (cfg.nb == 0) {
    set_error("stem forward: no configuration (K=%u c_out=%u)", K, c_cols);
    return MEB200_ERR_UNSUPPORTED;
  }
  const uint32_t c_virt = tc::stem_virtual_channels(K);
  TaParams p{};
  p.A = A4; p.nbr = nbr; p.out = out;
  p.c_red = c_virt; p.c_cols = c_cols; p.K = K; p.n_rows = n_rows;
  p.wb_row0 = 0; p.wb_rows_per_k = c_cols;
  p.n_super = cfg.n_super; p.R = cfg.R; p.acc_sets = cfg.acc_sets; p.nA = cfg.nA; p.nB = cfg.nB;
  p.n_macro = cfg.n_macro; p.b_stage_bytes = cfg.b_stage_bytes; p.b_sub_bytes = cfg.b_sub_bytes;
  p.a_col0 = cfg.a_col0; p.out_f32 = out_dtype == MEB200_F32; p.out_ld = c_cols;
  p.rs = cfg.rs;
  p.Kg = 1;
  CUtensorMap mB;
  memset(&mB, 0, sizeof(mB));
  int rc = make_map_2d(&mB, dtype, Wv, (uint64_t)c_cols, c_virt, (uint32_t)cfg.bk, c_cols);
  if (rc != MEB200_OK) return rc;
  const unsigned grid = p.n_super < (uint32_t)num_sms() ? p.n_super : (unsigned)num_sms();
  if (dtype == MEB200_BF16)
    return launch_ts<__nv_bfloat16, (int)tc::kStemNb, 64, 0>(mB, p, cfg.smem_bytes, grid, stream);
  return launch_ts<__half, (int)tc::kStemNb, 64, 0>(mB, p, cfg.smem_bytes, grid, stream);
}

bool conv_tc_supported(int dtype, uint32_t c_reduce, uint32_t c_cols) {
  if (dtype != MEB200_BF16 && dtype != MEB200_F16) return false;
  if (pick_bk(c_reduce) == 0) return false;
  return c_cols % 16 == 0 && c_cols >= 16 && c_cols <= 1024;
}

int conv_forward_tc(const void *A, int dtype, uint32_t n_a, uint32_t c_reduce, const void *W,
                    uint32_t K, uint32_t c_cols, bool dgrad, const int32_t *nbr, uint32_t n_rows,
                    void *out, int out_dtype, void *workspace, cudaStream_t stream,
                    const void *Wperm) {
  if (n_rows == 0) return MEB200_OK;
  MEB_CHECK_ARG(conv_tc_supported(dtype, c_reduce, c_cols), "shape not supported by tc path");
  const uint8_t *Wb = reinterpret_cast<const uint8_t *>(W);
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
    Wb = reinterpret_cast<const uint8_t *>(workspace);
  }
  // operand B is [K, c_cols, c_reduce]; split the column range in slices of <= 256
  const size_t out_esz = out_dtype == MEB200_F32 ? 4 : 2;
  const bool aligned16 = (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(Wperm) & 15) == 0;
  for (uint32_t n0 = 0; n0 < c_cols; n0 += 256) {
    uint32_t nc = c_cols - n0 < 256 ? c_cols - n0 : 256;
    if (Wperm != nullptr && aligned16 && c_reduce % 32 == 0 && ts_enabled()) {
      // operand A in tensor memory: needs the weights packed in ta_perm order
      int rc = launch_ta_slice(A, dtype, n_a, c_reduce, Wperm, c_cols, n0, K, nc, nbr, n_rows,
                               reinterpret_cast<uint8_t *>(out) + (size_t)n0 * out_esz, c_cols,
                               out_dtype, stream);
      if (rc == MEB200_OK) continue;
      if (rc != MEB200_ERR_UNSUPPORTED) return rc;
    }
    int rc = launch_tc_slice(A, dtype, n_a, c_reduce, Wb, c_cols, n0, K, nc, nbr, n_rows,
                             reinterpret_cast<uint8_t *>(out) + (size_t)n0 * out_esz, c_cols,
                             out_dtype, stream);
    if (rc != MEB200_OK) return rc;
  }
  return MEB200_OK;
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

struct WgParams {
  const void *in;        // [n_in, c_in]
  const void *gout;      // [n_out, c_out]
  const int32_t *nbr;    // [K, n_out]
  float *dW;             // [K, c_in, c_out]
  uint32_t c_in, c_out, K, n_out;
  uint32_t G, n_kgroups, mt_cta, n_mtgroups, rows_per_split, n_splits;
  uint32_t blkA, blkB;   // 64-channel blocks per A / B stage
  uint32_t nA, nB, a_stage_bytes, b_stage_bytes;
  uint32_t lag;
};

// The round-1 dense kernel: the fallback when the caller has no compacted pair lists.
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
  const uint32_t k0 = tc::kgroup_begin(kg, p.K, p.n_kgroups);
  const uint32_t G = tc::kgroup_begin(kg + 1, p.K, p.n_kgroups) - k0;   // <= p.G
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
    uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
    // neighbour rows for (chunk c, offset kk), fetched ahead of use
    auto load_idx = [&](uint32_t c, uint32_t kk, int32_t (&dst)[kWgRows / 16]) {
      const int32_t *nbr_k = p.nbr + (size_t)(k0 + kk) * p.n_out;
#pragma unroll
      for (int it = 0; it < kWgRows / 16; ++it) {
        const uint32_t o = r_begin + c * kWgRows + it * 16 + warp * 4 + sub_row;
        dst[it] = (o < r_end) ? __ldg(nbr_k + o) : -1;
      }
    };
    // one work item: (re)use ring slots, issue the copies of stage (c, kk), publish them
    auto issue_stage = [&](uint32_t c, uint32_t kk, const int32_t (&idx_cur)[kWgRows / 16]) {
        const uint32_t row0 = r_begin + c * kWgRows;
        mbar_wait(emptyA(sa), pha ^ 1u);
        if (kk == 0) {
          mbar_wait(emptyB(sb), phb ^ 1u);
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
        const uint32_t dstA = sA + sa * p.a_stage_bytes;
        for (uint32_t blk = 0; blk < p.blkA; ++blk) {
          const uint32_t chn = ch0 + blk * 64 + j8 * 8;
#pragma unroll
          for (int it = 0; it < kWgRows / 16; ++it) {
            const uint32_t row = it * 16 + warp * 4 + sub_row;
            const bool ok = (idx_cur[it] >= 0) && (chn < p.c_in);
            const T *src = in + (ok ? ((size_t)idx_cur[it] * p.c_in + chn) : 0);
            cp_async16(dstA + blk * kBlkBytes + row * 128 + ((j8 ^ (row & 7u)) << 4), src,
                       ok ? 16u : 0u);
          }
        }
        if (kk == 0) {
          cp_async_mbar_arrive(fullB(sb));
          if (++sb == p.nB) { sb = 0; phb ^= 1u; }
        }
        cp_async_mbar_arrive(fullA(sa));
        if (++sa == p.nA) { sa = 0; pha ^= 1u; }
    };
    int32_t idx_cur[kWgRows / 16], idx_nxt[kWgRows / 16];
    if (n_chunks > 0) load_idx(0, 0, idx_cur);
    for (uint32_t c = 0; c < n_chunks; ++c) {
      for (uint32_t kk = 0; kk < G; ++kk) {
        {
          uint32_t nc = c, nkk = kk + 1;
          if (nkk == G) { nkk = 0; nc = c + 1; }
          if (nc < n_chunks) load_idx(nc, nkk, idx_nxt);
        }
        issue_stage(c, kk, idx_cur);
#pragma unroll
        for (int it = 0; it < kWgRows / 16; ++it) idx_cur[it] = idx_nxt[it];
      }
    }
    cp_async_wait<0>();
  } else if (warp < kProducerWarps + kEpilogueWarps) {
    const uint32_t q = warp & 3u;
    mbar_wait_park(tfull, 0);              // parked for the whole main loop
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
    uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
    for (uint32_t c = 0; c < n_chunks; ++c) {
      mbar_wait(fullB(sb), phb);
      const uint32_t bbase = sB + sb * p.b_stage_bytes;
      for (uint32_t kk = 0; kk < G; ++kk) {
        mbar_wait(fullA(sa), pha);
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
        if (++sa == p.nA) { sa = 0; pha ^= 1u; }
      }
      umma_commit(emptyB(sb));
      if (++sb == p.nB) { sb = 0; phb ^= 1u; }
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

// =====================================================================================
// wgrad over COMPACTED PAIR LISTS (k_wgrad_pairs, default when the caller provides the lists).
//
// The dense kernel above spends 69 % of its stages' rows on zero rows (8.5 of 27 neighbours
// exist on the bench clouds) and its four producer warps are issue bound (profiles/r2_notes.md).
// Here the reduction runs over the valid (input row, output row) pairs of one offset at a time
// (meb200_kernel_map_pairs: per-offset segments padded to whole stages), both operands gathered
// row by row with cp.async:
//   stage    = 64 pairs: A = 64 gathered input rows (this CTA's <= 256-channel slice),
//              B = the 64 matching dOut rows; MN-major 128B-swizzled blocks as above
//   CTA      = (channel slice, contiguous range of the global stage list); when the range
//              crosses an offset boundary the accumulator is flushed (red.global.add into
//              dW[k]) and the next offset starts on the other accumulator set
//   warps    = 8 producers (8 rows of every stage each; per-thread copy slots are fixed, so a
//              stage costs one shuffle + one address + one cp.async per 16 bytes), 4 epilogue
//              warps, 1 MMA warp
// No zero fill (channels past c_in in the last A block feed accumulator rows nobody reads),
// no wasted MMAs except the <= 63 pad pairs at the end of each offset.
// =====================================================================================
constexpr int kWpRows = 64;
constexpr int kWpGroups = 3;                // producer groups: group g owns stages n = g (mod 3)
constexpr int kWpProducerWarps = 4 * kWpGroups;   // 4 warps per group, 16 pairs of a stage each
constexpr int kWpRowsPerWarp = 16;
constexpr int kWpThreads = (kWpProducerWarps + kEpilogueWarps + 1) * 32;
constexpr int kWpPf = 4;                    // pair indices are copied this many of the group's stages ahead
constexpr uint32_t kWpMaxSeg = 2047;        // (row chunks) x K segments a launch can walk

struct WpParams {
  const void *in;        // [n_in, c_in]
  const void *gout;      // [n_out, c_out]
  const int32_t *pin, *pout, *seg_start;   // padded pair lists, seg_start[n_chunks*K + 1]
  float *dW;             // [K, c_in, c_out]
  uint32_t c_in, c_out, K, n_chunks;
  uint32_t mt_cta, n_mtgroups, n_splits;
  uint32_t blkB, n_stage, a_bytes, stage_bytes, acc_sets;
};

// This CTA's share of the stage list.  The list is physically (row chunk, offset)-ordered; a CTA
// owns a contiguous range [a, b) of the OFFSET-major cumulative stage count (so it touches one or
// two offsets and flushes its accumulator once per offset), and takes from EVERY chunk's segment
// of such an offset the same fraction of the stages.  All CTAs therefore sweep the chunks together
// (one chunk's rows of both feature matrices stay L2-resident) without flushing per chunk.
// Every role iterates identically.
struct WpWalk {
  const int32_t *seg;      // [C*K + 1], in stages (shared memory)
  uint32_t K, C, a, b;
  uint32_t k, c, Sk, Tk, x0, x1, s, s_end;
  bool fin;
  __device__ __forceinline__ uint32_t seg_len(uint32_t cc, uint32_t kk) const {
    return (uint32_t)(seg[cc * K + kk + 1] - seg[cc * K + kk]);
  }
  __device__ __forceinline__ uint32_t total_of(uint32_t kk) const {
    uint32_t t = 0;
    for (uint32_t cc = 0; cc < C; ++cc) t += seg_len(cc, kk);
    return t;
  }
  __device__ __forceinline__ bool open_c() {      // first non-empty run of offset k at chunk >= c
    for (; c < C; ++c) {
      const uint32_t n = seg_len(c, k), lo = (uint32_t)seg[c * K + k];
      s = lo + (uint32_t)((uint64_t)n * x0 / Tk);
      s_end = lo + (uint32_t)((uint64_t)n * x1 / Tk);
      if (s < s_end) return true;
    }
    return false;
  }
  __device__ __forceinline__ void open_k() {      // first offset >= k this CTA has stages of
    for (; k < K; ++k) {
      Tk = total_of(k);
      if (Sk + Tk > a && Sk < b && Tk > 0) {
        x0 = (a > Sk ? a : Sk) - Sk;
        x1 = (b < Sk + Tk ? b : Sk + Tk) - Sk;
        c = 0;
        if (x0 < x1 && open_c()) return;
      }
      Sk += Tk;
      if (Sk >= b) break;
    }
    fin = true;
  }
  __device__ __forceinline__ void init(const int32_t *seg_, uint32_t K_, uint32_t C_,
                                       uint32_t split, uint32_t n_splits) {
    seg = seg_; K = K_; C = C_;
    const uint32_t total = (uint32_t)seg[C * K];
    a = (uint32_t)((uint64_t)total * split / n_splits);
    b = (uint32_t)((uint64_t)total * (split + 1) / n_splits);
    k = 0; c = 0; Sk = 0; Tk = 0; x0 = x1 = 0; s = s_end = 0; fin = a >= b;
    if (!fin) open_k();
  }
  __device__ __forceinline__ bool done() const { return fin; }
  __device__ __forceinline__ void after_run() {   // s reached s_end
    ++c;
    if (!open_c()) { Sk += Tk; ++k; if (Sk >= b) fin = true; else open_k(); }
  }
  // next run of stages inside one (chunk, offset) segment: [rb, re) of offset kk
  __device__ __forceinline__ void next_run(uint32_t &rb, uint32_t &re, uint32_t &kk) {
    rb = s; re = s_end; kk = k;
    s = s_end;
    after_run();
  }
  // single-stage stepping for the producers
  __device__ __forceinline__ uint32_t next_stage() {
    const uint32_t cur = s;
    if (++s >= s_end) after_run();
    return cur;
  }
  __device__ __forceinline__ void skip(uint32_t n) {      // pass over n stages
    while (n > 0 && !fin) {
      const uint32_t left = s_end - s;
      if (n < left) { s += n; return; }
      n -= left;
      s = s_end;
      after_run();
    }
  }
};

// S = copy slots per thread per operand per 8 rows = ceil(channels / 32), a template parameter so
// that the slot tables live in registers and no dead slot is ever stepped over (with 8 generic
// slots the loop was 672 instructions per stage and warp, its tables spilled to local memory).
// RW = pairs per stage (64, or 32 for wide layers: with 64-row stages a 256 + 256 channel layer
// fits only 3 stages = one per producer group, and a group then runs issue -> landing -> MMA
// strictly in sequence: 14 k cycles per stage on block4 256->256, r2_run15).
template <typename T, int S, int RW>
__global__ void __launch_bounds__(kWpThreads, 1) k_wgrad_pairs(const WpParams p) {
  constexpr uint32_t kBlk = RW * 128u;            // one 64-channel block of a stage
  constexpr uint32_t kHalves = RW / 32;           // 8-row halves per producer warp
  constexpr uint32_t kRowsPerWarp = RW / 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s0 = smem_u32(smem);
  uint8_t *tail = smem + p.n_stage * p.stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
  const uint32_t bar0 = smem_u32(bars);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 * kMaxStages + 4);
  int32_t *iring_all = reinterpret_cast<int32_t *>(tail + tc::kBarBytes);   // [warps][kWpPf][32]
  int32_t *seg = iring_all + kWpProducerWarps * kWpPf * 32;                  // [n_seg + 1], in STAGES

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t mtg = blockIdx.x % p.n_mtgroups, split = blockIdx.x / p.n_mtgroups;
  const uint32_t ch0 = mtg * p.mt_cta * 128;
  const uint32_t ch_n = min(p.c_in - ch0, p.mt_cta * 128);        // channels of this CTA's slice
  const uint32_t mt = (ch_n + 127) / 128;
  const uint32_t n_seg = p.n_chunks * p.K;

  for (uint32_t i = threadIdx.x; i <= n_seg; i += kWpThreads) seg[i] = __ldg(p.seg_start + i) / RW;
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < p.n_stage; ++i) {
      mbar_init(bar_fullA(bar0, i), 4 * 32);       // the four warps of the owning group
      mbar_init(bar_emptyA(bar0, i), 1);
    }
    for (uint32_t i = 0; i < 2; ++i) { mbar_init(bar_tfull(bar0, i), 1); mbar_init(bar_tempty(bar0, i), kEpilogueWarps * 32); }
    mbar_fence_init();
  }
  if (warp == kWpProducerWarps + kEpilogueWarps) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWpProducerWarps) {
    // =============================== PRODUCERS ======================================
    // The producers are bound by the instruction latency of a single warp (~10 cycles per
    // instruction, ~150 instructions per stage): with all warps filling every stage in lockstep a
    // stage cost ~1500 cycles whatever the memory system did (r2_run5/7).  So the stages are dealt
    // to kWpGroups groups of 4 warps; a group fills its stage alone (16 pairs per warp, as two
    // halves of 8 rows over the fixed copy slots) while the other groups fill theirs.
    const uint32_t grp = warp >> 2, qw = warp & 3u;
    const T *in = reinterpret_cast<const T *>(p.in) + ch0;
    const T *gout = reinterpret_cast<const T *>(p.gout);
    const uint32_t cprA = ch_n / 8, cprB = p.c_out / 8;      // 16-byte chunks per row
    const uint32_t nA = (8 * cprA + 31) / 32, nB = (8 * cprB + 31) / 32;   // slots per thread per 8 rows
    // fixed copy slots: slot s of this lane copies chunk `c % cpr` of row `c / cpr` (of 8 rows),
    // c = lane + 32 s; destination offsets never change (the second half adds 8 rows = 1024 B)
    uint32_t a_dst[S], b_dst[S], a_meta[S], b_meta[S];
#pragma unroll
    for (int sl = 0; sl < S; ++sl) {
      const uint32_t c = lane + 32u * sl;
      {
        const uint32_t rl = c / cprA, chunk = c - rl * cprA;
        const bool ok = (uint32_t)sl < nA && rl < 8;
        a_meta[sl] = ok ? (rl | (chunk << 3)) : 0xffffffffu;
        a_dst[sl] = (chunk >> 3) * kBlk + (qw * kRowsPerWarp + rl) * 128 + (((chunk & 7u) ^ rl) << 4);
      }
      {
        const uint32_t rl = c / cprB, chunk = c - rl * cprB;
        const bool ok = (uint32_t)sl < nB && rl < 8;
        b_meta[sl] = ok ? (rl | (chunk << 3)) : 0xffffffffu;
        b_dst[sl] = p.a_bytes + (chunk >> 3) * kBlk + (qw * kRowsPerWarp + rl) * 128 +
                    (((chunk & 7u) ^ rl) << 4);
      }
    }
    // The pair lists stream from HBM: the 16 + 16 row indices of the warp's pairs of a stage are
    // copied kWpPf of the group's stages ahead by cp.async (lanes 0-15 input rows, 16-31 output
    // rows) into a per-warp ring in shared memory, one commit group per stage (a register ring
    // would make every stage wait for its newest load).
    const int32_t *list = lane < 16 ? p.pin : p.pout;
    const uint32_t li = qw * kRowsPerWarp + (lane & 15u);
    const bool li_ok = (lane & 15u) < kRowsPerWarp;
    const uint32_t iring = smem_u32(iring_all + warp * kWpPf * 32);
    WpWalk wk, wpf;
    wk.init(seg, p.K, p.n_chunks, split, p.n_splits);
    wk.skip(grp);
    wpf.init(seg, p.K, p.n_chunks, split, p.n_splits);
    wpf.skip(grp);
    // (a macro, not a lambda: a by-reference capture would put the walkers in local memory)
#define MEB_WP_COPY_IDX(nn)                                                                       \
    do {                                                                                          \
      if (!wpf.done()) {                                                                          \
        const uint32_t stage__ = wpf.s;                                                           \
        wpf.skip(kWpGroups);                                                                      \
        if (li_ok)                                                                                \
          cp_async4(iring + (((nn) % kWpPf) * 32 + lane) * 4, list + (size_t)stage__ * RW + li, 4u); \
      }                                                                                           \
      cp_async_commit();                                                                          \
    } while (0)
#pragma unroll 1
    for (uint32_t d = 0; d < (uint32_t)kWpPf; ++d) MEB_WP_COPY_IDX(d);
    // the group's stage n uses ring slot (grp + 3 n) % n_stage; n_stage is a multiple of 3, so a
    // slot always belongs to the same group and a one-bit parity wait on "empty" is exact
    uint32_t slot = grp, ph = 0, n = 0;
    while (!wk.done()) {
      wk.skip(kWpGroups);
      cp_async_wait<kWpPf - 1>();      // this stage's indices (copied kWpPf stages ago) landed
      if (lane == 0) mbar_wait_park(bar_emptyA(bar0, slot), ph ^ 1u);
      __syncwarp();
      const uint32_t base = s0 + slot * p.stage_bytes;
      const uint32_t irow = iring + (n % kWpPf) * 128;
#pragma unroll
      for (int h = 0; h < (int)kHalves; ++h) {
#pragma unroll
        for (int sl = 0; sl < S; ++sl) {
          if ((uint32_t)sl < nA) {
            const uint32_t m = a_meta[sl];
            if (m != 0xffffffffu) {
              int32_t row;
              asm volatile("ld.shared.s32 %0, [%1];" : "=r"(row) : "r"(irow + (h * 8 + (m & 7u)) * 4));
              const T *src = in + (row < 0 ? 0 : (size_t)row * p.c_in + (m >> 3) * 8);
              cp_async16(base + a_dst[sl] + h * 1024, src, row < 0 ? 0u : 16u);
            }
          }
        }
#pragma unroll
        for (int sl = 0; sl < S; ++sl) {
          if ((uint32_t)sl < nB) {
            const uint32_t m = b_meta[sl];
            if (m != 0xffffffffu) {
              int32_t row;
              asm volatile("ld.shared.s32 %0, [%1];" : "=r"(row) : "r"(irow + (16 + h * 8 + (m & 7u)) * 4));
              const T *src = gout + (row < 0 ? 0 : (size_t)row * p.c_out + (m >> 3) * 8);
              cp_async16(base + b_dst[sl] + h * 1024, src, row < 0 ? 0u : 16u);
            }
          }
        }
      }
      cp_async_mbar_arrive(bar_fullA(bar0, slot));
      slot += kWpGroups;
      if (slot >= p.n_stage) { slot -= p.n_stage; ph ^= 1u; }
      __syncwarp();                    // ring row n % kWpPf fully read: reuse it
      MEB_WP_COPY_IDX(n + kWpPf);
      ++n;
    }
#undef MEB_WP_COPY_IDX
    cp_async_wait<0>();
  } else if (warp < kWpProducerWarps + kEpilogueWarps) {
    // =============================== EPILOGUE =======================================
    // one flush per offset this CTA has stages of (the walker yields its runs offset by offset)
    const uint32_t q = warp & 3u;
    uint32_t acc = 0, par = 0;
    WpWalk wk;
    wk.init(seg, p.K, p.n_chunks, split, p.n_splits);
    while (!wk.done()) {
      uint32_t rb, re, k, k2;
      wk.next_run(rb, re, k);
      while (!wk.done() && wk.k == k) wk.next_run(rb, re, k2);     // skip this offset's other runs
      mbar_wait_park(bar_tfull(bar0, acc), par);
      tc_fence_after();
      for (uint32_t t = 0; t < mt; ++t) {
        const uint32_t ci = ch0 + t * 128 + q * 32 + lane;
        const uint32_t col0 = (acc * p.mt_cta + t) * p.c_out;
        float *dst_row = p.dW + ((size_t)k * p.c_in + ci) * p.c_out;
        for (uint32_t c = 0; c < p.c_out; c += 16) {
          uint32_t v[16];
          tmem_ld16(tmem_base + ((q * 32u) << 16) + col0 + c, v);
          tmem_ld_wait();
          if (ci < p.c_in) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)       // 16-byte vector reductions (sm_90+)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + c + i),
                           "f"(__uint_as_float(v[i])), "f"(__uint_as_float(v[i + 1])),
                           "f"(__uint_as_float(v[i + 2])), "f"(__uint_as_float(v[i + 3]))
                           : "memory");
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty(bar0, acc));
      if (++acc == p.acc_sets) { acc = 0; par ^= 1u; }
    }
  } else {
    // =============================== MMA ISSUER =====================================
    const uint32_t idesc = umma_idesc_f16(Fmt<T>::v, kTileM, p.c_out, 1, 1);
    uint32_t slot = 0, ph = 0, acc = 0, par = 0;
    WpWalk wk;
    wk.init(seg, p.K, p.n_chunks, split, p.n_splits);
    while (!wk.done()) {
      const uint32_t k_cur = wk.k;
      mbar_wait_park(bar_tempty(bar0, acc), par ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + acc * p.mt_cta * p.c_out;
      bool first = true;
      while (!wk.done() && wk.k == k_cur) {          // all chunks' runs of this offset
        uint32_t rb, re, k;
        wk.next_run(rb, re, k);
        for (uint32_t st = rb; st < re; ++st) {
          mbar_wait_park(bar_fullA(bar0, slot), ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t abase = s0 + slot * p.stage_bytes, bbase = abase + p.a_bytes;
            for (uint32_t t = 0; t < mt; ++t) {
#pragma unroll
              for (int j = 0; j < RW / 16; ++j) {
                const uint64_t da = umma_desc(abase + t * 2 * kBlk + j * 2048, kBlk, 1024, 2);
                const uint64_t db = umma_desc(bbase + j * 2048, kBlk, 1024, 2);
                umma_f16(d0 + t * p.c_out, da, db, idesc, (!first || j != 0) ? 1u : 0u);
              }
            }
            umma_commit(bar_emptyA(bar0, slot));
          }
          __syncwarp();
          first = false;
          if (++slot == p.n_stage) { slot = 0; ph ^= 1u; }
        }
      }
      if (elect_one()) umma_commit(bar_tfull(bar0, acc));
      __syncwarp();
      if (++acc == p.acc_sets) { acc = 0; par ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWpProducerWarps + kEpilogueWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

bool conv_wgrad_pairs_supported(int dtype, uint32_t c_in, uint32_t K, uint32_t c_out) {
  if (dtype != MEB200_BF16 && dtype != MEB200_F16) return false;
  return c_in % 8 == 0 && c_in >= 16 && c_out % 16 == 0 && c_out >= 16 && c_out <= 256 &&
         K <= kWpMaxSeg;
}

int conv_wgrad_pairs(const void *in, const void *grad_out, int dtype, uint32_t c_in, uint32_t K,
                     uint32_t c_out, const int32_t *pairs_in, const int32_t *pairs_out,
                     const int32_t *seg_start, uint32_t n_chunks, uint32_t n_out,
                     float *grad_weight, cudaStream_t stream) {
  MEB_CUDA(cudaMemsetAsync(grad_weight, 0, (size_t)K * c_in * c_out * sizeof(float), stream));
  if (n_out == 0 || K == 0) return MEB200_OK;
  const tc::WpCfg cfg = tc::wgrad_pairs_config(c_in, c_out, K, n_chunks, n_out, (uint32_t)num_sms());
  if (cfg.mt_cta == 0) {
    set_error("wgrad pairs: no configuration fits (c_in=%u c_out=%u K=%u)", c_in, c_out, K);
    return MEB200_ERR_UNSUPPORTED;
  }
  WpParams p{};
  p.in = in; p.gout = grad_out; p.pin = pairs_in; p.pout = pairs_out; p.seg_start = seg_start;
  p.dW = grad_weight; p.c_in = c_in; p.c_out = c_out; p.K = K; p.n_chunks = n_chunks;
  p.mt_cta = cfg.mt_cta; p.n_mtgroups = cfg.n_mtgroups; p.n_splits = cfg.n_splits;
  p.blkB = cfg.blkB; p.n_stage = cfg.n_stage; p.a_bytes = cfg.a_bytes;
  p.stage_bytes = cfg.stage_bytes; p.acc_sets = cfg.acc_sets;
  const unsigned grid = cfg.n_mtgroups * cfg.n_splits;
  // slots per thread: the wider of the two operands' rows (A: this CTA's channel slice)
  const uint32_t ch_a = c_in < cfg.mt_cta * 128 ? c_in : cfg.mt_cta * 128;
  const uint32_t slots = ((ch_a > c_out ? ch_a : c_out) + 31) / 32;
#define MEB_WP(SV, RWV)                                                                         \
  do {                                                                                          \
    if (dtype == MEB200_BF16) {                                                                 \
      MEB_BIG_SMEM((k_wgrad_pairs<__nv_bfloat16, SV, RWV>));                                    \
      k_wgrad_pairs<__nv_bfloat16, SV, RWV><<<grid, kWpThreads, cfg.smem_bytes, stream>>>(p);   \
    } else {                                                                                    \
      MEB_BIG_SMEM((k_wgrad_pairs<__half, SV, RWV>));                                           \
      k_wgrad_pairs<__half, SV, RWV><<<grid, kWpThreads, cfg.smem_bytes, stream>>>(p);          \
    }                                                                                           \
  } while (0)
  if (cfg.rw == 32) {           // wide layers only (the narrow ones fit >= 6 stages of 64 pairs)
    if (slots <= 4) MEB_WP(4, 32);
    else if (slots <= 6) MEB_WP(6, 32);
    else MEB_WP(8, 32);
  } else if (slots <= 1) MEB_WP(1, 64);
  else if (slots == 2) MEB_WP(2, 64);
  else if (slots == 3) MEB_WP(3, 64);
  else if (slots == 4) MEB_WP(4, 64);
  else if (slots <= 6) MEB_WP(6, 64);
  else MEB_WP(8, 64);
#undef MEB_WP
  count_tc_launch();
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

// =====================================================================================
// wgrad of a network STEM (rows of 4 channels, see k_conv_ts STEM): the reduction runs over ALL
// output rows (dense, output-stationary on dW), the M axis is (offset, channel):
//   dWv[4 k + c][n] = sum_r in4[nbr[k][r]][c] * dOut[r][n]        k < Kp = 16 ceil(K / 16)
// A stage = 32 output rows: A = [32 rows x 4 Kp virtual channels] gathered 8 bytes per
// (row, offset) straight into MN-major 128B-swizzled blocks (missing neighbours zero-filled by
// the copy engine), B = the 32 dOut rows.  Every CTA keeps ONE accumulator set for its whole row
// share ([4 Kp x c_out] fp32 in tensor memory) and flushes it once.  Roles as in k_wgrad_pairs:
// 3 producer groups of 4 warps (a group owns every third stage; lane = row, so the table reads
// are coalesced; the indices of the next stage are loaded while this stage's copies fly),
// 4 epilogue warps, 1 MMA warp.  Replaces k_conv_small_cin_wgrad (1.0 ms on 3 -> 32, K = 125,
// 800k rows) on the bf16/fp16 path.
// =====================================================================================
constexpr int kWsRows = 32;
constexpr uint32_t kWsBlk = kWsRows * 128u;
constexpr int kWsGroups = 3;
constexpr int kWsProducerWarps = 4 * kWsGroups;
constexpr int kWsThreads = (kWsProducerWarps + kEpilogueWarps + 1) * 32;
constexpr int kWsMaxOffsetsPerThread = 32;     // Kp / 4 <= 32  <=>  K <= 128

struct WsParams {
  const void *in4;       // [n_in, 4]
  const void *gout;      // [n_rows, c_out]
  const int32_t *nbr;    // [K, n_rows]
  float *dWv;            // [4 Kp, c_out], zero-filled
  uint32_t K, Kp, n_rows, c_out;
  uint32_t mt, n_stage, a_bytes, stage_bytes, n_stages_total;
};

template <typename T>
__global__ void __launch_bounds__(kWsThreads, 1) k_wgrad_stem(const WsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s0 = smem_u32(smem);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + p.n_stage * p.stage_bytes);
  const uint32_t bar0 = smem_u32(bars);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 * kMaxStages + 4);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t n_my = p.n_stages_total > blockIdx.x
                            ? (p.n_stages_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < p.n_stage; ++i) {
      mbar_init(bar_fullA(bar0, i), 4 * 32);
      mbar_init(bar_emptyA(bar0, i), 1);
    }
    mbar_init(bar_tfull(bar0, 0), 1);
    mbar_fence_init();
  }
  if (warp == kWsProducerWarps + kEpilogueWarps) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kWsProducerWarps) {
    // =============================== PRODUCERS ======================================
    const uint32_t grp = warp >> 2, qw = warp & 3u;
    const T *in4 = reinterpret_cast<const T *>(p.in4);
    const T *gout = reinterpret_cast<const T *>(p.gout);
    const uint32_t n_off = p.Kp / 4;               // offsets per thread: k = qw + 4 i
    const uint32_t cprB = p.c_out / 8;             // 16-byte chunks per dOut row
    int32_t idx[kWsMaxOffsetsPerThread];
    auto load_idx = [&](uint32_t stage) {
      const uint32_t row = stage * kWsRows + lane;
#pragma unroll
      for (int i = 0; i < kWsMaxOffsetsPerThread; ++i) {
        const uint32_t k = qw + 4u * i;
        idx[i] = ((uint32_t)i < n_off && k < p.K && row < p.n_rows)
                     ? __ldg(p.nbr + (size_t)k * p.n_rows + row) : -1;
      }
    };
    uint32_t slot = grp, ph = 0;
    if (grp < n_my) load_idx(blockIdx.x + grp * gridDim.x);
    for (uint32_t i = grp; i < n_my; i += kWsGroups) {
      const uint32_t stage = blockIdx.x + i * gridDim.x;
      if (lane == 0) mbar_wait_park(bar_emptyA(bar0, slot), ph ^ 1u);
      __syncwarp();
      const uint32_t base = s0 + slot * p.stage_bytes;
      const uint32_t rowoff = lane * 128u, rx = lane & 7u;
#pragma unroll
      for (int j = 0; j < kWsMaxOffsetsPerThread; ++j) {
        if ((uint32_t)j < n_off) {
          const uint32_t k = qw + 4u * j;
          const int32_t v = idx[j];
          const uint32_t dst = base + (k >> 4) * kWsBlk + rowoff + ((((k & 15u) >> 1) ^ rx) << 4) + (k & 1u) * 8u;
          cp_async8(dst, in4 + (v < 0 ? 0 : (size_t)v * 4), v < 0 ? 0u : 8u);
        }
      }
      // the 32 dOut rows of the stage: chunk c of row r, (r, c) dealt over the group's threads
      for (uint32_t c = qw * 32 + lane; c < kWsRows * cprB; c += 128) {
        const uint32_t r = c / cprB, ch = c - r * cprB;
        const uint32_t row = stage * kWsRows + r;
        const uint32_t dst = base + p.a_bytes + (ch >> 3) * kWsBlk + r * 128u + (((ch & 7u) ^ (r & 7u)) << 4);
        const bool ok = row < p.n_rows;
        cp_async16(dst, gout + (ok ? (size_t)row * p.c_out + ch * 8 : 0), ok ? 16u : 0u);
      }
      cp_async_mbar_arrive(bar_fullA(bar0, slot));
      slot += kWsGroups;
      if (slot >= p.n_stage) { slot -= p.n_stage; ph ^= 1u; }
      if (i + kWsGroups < n_my) load_idx(blockIdx.x + (i + kWsGroups) * gridDim.x);
    }
    cp_async_wait<0>();
  } else if (warp < kWsProducerWarps + kEpilogueWarps) {
    // =============================== EPILOGUE: one flush ================================
    const uint32_t q = warp & 3u;
    if (n_my > 0) {
      mbar_wait_park(bar_tfull(bar0, 0), 0);
      tc_fence_after();
      for (uint32_t t = 0; t < p.mt; ++t) {
        const uint32_t v = t * 128 + q * 32 + lane;              // virtual channel = 4 k + c
        float *dst_row = p.dWv + (size_t)v * p.c_out;
        for (uint32_t c = 0; c < p.c_out; c += 16) {
          uint32_t r[16];
          tmem_ld16(tmem_base + ((q * 32u) << 16) + t * p.c_out + c, r);
          tmem_ld_wait();
          if (v < 4 * p.Kp) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_row + c + i),
                           "f"(__uint_as_float(r[i])), "f"(__uint_as_float(r[i + 1])),
                           "f"(__uint_as_float(r[i + 2])), "f"(__uint_as_float(r[i + 3]))
                           : "memory");
          }
        }
      }
      tc_fence_before();
    }
  } else {
    // =============================== MMA ISSUER =====================================
    const uint32_t idesc = umma_idesc_f16(Fmt<T>::v, kTileM, p.c_out, 1, 1);
    uint32_t slot = 0, ph = 0;
    for (uint32_t i = 0; i < n_my; ++i) {
      mbar_wait_park(bar_fullA(bar0, slot), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t abase = s0 + slot * p.stage_bytes, bbase = abase + p.a_bytes;
        for (uint32_t t = 0; t < p.mt; ++t) {
#pragma unroll
          for (int j = 0; j < kWsRows / 16; ++j) {
            const uint64_t da = umma_desc(abase + t * 2 * kWsBlk + j * 2048, kWsBlk, 1024, 2);
            const uint64_t db = umma_desc(bbase + j * 2048, kWsBlk, 1024, 2);
            umma_f16(tmem_base + t * p.c_out, da, db, idesc, (i != 0 || j != 0) ? 1u : 0u);
          }
        }
        umma_commit(bar_emptyA(bar0, slot));
      }
      __syncwarp();
      if (++slot == p.n_stage) { slot = 0; ph ^= 1u; }
    }
    if (n_my > 0 && elect_one()) umma_commit(bar_tfull(bar0, 0));
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWsProducerWarps + kEpilogueWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

bool conv_stem_wgrad_tc_supported(int dtype, uint32_t K, uint32_t c_out) {
  if (dtype != MEB200_BF16 && dtype != MEB200_F16) return false;
  const uint32_t Kp = (K + 15) / 16 * 16;
  const uint32_t mt = (4 * Kp + 127) / 128;
  return K >= 1 && Kp / 4 <= (uint32_t)kWsMaxOffsetsPerThread && c_out % 16 == 0 && c_out >= 16 &&
         c_out <= 256 && mt * c_out <= kTmemCols && encode_tiled_fn() != nullptr;
}

// dWv: [4 * 16 ceil(K / 16), c_out] fp32 (virtual channel 4 k + c), zero-filled here.
int conv_stem_wgrad_tc(const void *in4, const void *grad_out, int dtype, uint32_t K,
                       uint32_t c_out, const int32_t *nbr, uint32_t n_rows, float *dWv,
                       cudaStream_t stream) {
  MEB_CHECK_ARG(conv_stem_wgrad_tc_supported(dtype, K, c_out), "stem wgrad: K=%u c_out=%u", K, c_out);
  WsParams p{};
  p.in4 = in4; p.gout = grad_out; p.nbr = nbr; p.dWv = dWv;
  p.K = K; p.Kp = (K + 15) / 16 * 16; p.n_rows = n_rows; p.c_out = c_out;
  MEB_CUDA(cudaMemsetAsync(dWv, 0, (size_t)4 * p.Kp * c_out * sizeof(float), stream));
  if (n_rows == 0) return MEB200_OK;
  p.mt = (4 * p.Kp + 127) / 128;
  p.a_bytes = (4 * p.Kp / 64) * kWsBlk;
  if (p.a_bytes < 2 * kWsBlk) p.a_bytes = 2 * kWsBlk;      // an m-tile always reads two blocks
  p.stage_bytes = p.a_bytes + ((c_out + 63) / 64) * kWsBlk;
  const uint32_t budget = tc::kSmemBudget - 1024 - tc::kBarBytes - 64;
  p.n_stage = budget / p.stage_bytes;
  if (p.n_stage > 12) p.n_stage = 12;
  p.n_stage = p.n_stage / kWsGroups * kWsGroups;
  MEB_CHECK_ARG(p.n_stage >= (uint32_t)kWsGroups, "stem wgrad: stage of %u bytes does not fit", p.stage_bytes);
  p.n_stages_total = cdiv(n_rows, (uint32_t)kWsRows);
  const size_t smem_bytes = 1024 + (size_t)p.n_stage * p.stage_bytes + tc::kBarBytes + 64;
  const unsigned grid = p.n_stages_total < (uint32_t)num_sms() ? p.n_stages_total : (unsigned)num_sms();
  if (dtype == MEB200_BF16) {
    MEB_BIG_SMEM((k_wgrad_stem<__nv_bfloat16>));
    k_wgrad_stem<__nv_bfloat16><<<grid, kWsThreads, smem_bytes, stream>>>(p);
  } else {
    MEB_BIG_SMEM((k_wgrad_stem<__half>));
    k_wgrad_stem<__half><<<grid, kWsThreads, smem_bytes, stream>>>(p);
  }
  count_tc_launch();
  MEB_LAUNCH_OK();
  return MEB200_OK;
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
  const tc::WgCfg cfg = tc::wgrad_config(c_in, c_out, K, n_out, (uint32_t)num_sms());
  if (cfg.mt_cta == 0) {
    set_error("wgrad tc: no pipeline configuration fits (c_in=%u c_out=%u K=%u)", c_in, c_out, K);
    return MEB200_ERR_UNSUPPORTED;
  }
  p.mt_cta = cfg.mt_cta; p.n_mtgroups = cfg.n_mtgroups; p.G = cfg.G; p.n_kgroups = cfg.n_kgroups;
  p.blkA = cfg.blkA; p.blkB = cfg.blkB;
  p.a_stage_bytes = cfg.a_stage_bytes; p.b_stage_bytes = cfg.b_stage_bytes;
  p.lag = cfg.pipe.lag; p.nA = cfg.pipe.nA; p.nB = cfg.pipe.nB;
  p.rows_per_split = cfg.rows_per_split; p.n_splits = cfg.n_splits;
  size_t smem_bytes = cfg.smem_bytes;
  unsigned grid = cfg.n_kgroups * cfg.n_mtgroups * cfg.n_splits;
  int rc = MEB200_OK;
  if (dtype == MEB200_BF16) {
    MEB_BIG_SMEM(k_wgrad_tc<__nv_bfloat16>);
    k_wgrad_tc<__nv_bfloat16><<<grid, kThreads, smem_bytes, stream>>>(p);
  } else {
    MEB_BIG_SMEM(k_wgrad_tc<__half>);
    k_wgrad_tc<__half><<<grid, kThreads, smem_bytes, stream>>>(p);
  }
  if (rc != MEB200_OK) return rc;
  count_tc_launch();
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

}  // namespace meb200
