// One-shot all-reduce of a small fp64 vector over NVLink peer memory — the statistics exchange
// of synchronised batch norm (reference: torch.nn.SyncBatchNorm behind
// MinkowskiEngine/MinkowskiNormalization.py:101-192; NCCL all-reduce per layer there).
//
// A SyncBN exchange moves 2C+1 doubles (<= 4 KB) per layer, forward and backward: ~140 latency-
// bound NCCL launches per MinkUNet34C step, each with two cross-stream hand-offs (measured: 83 %
// weak-scaling efficiency at 2 GPUs, profiles/r1_notes.md).  Here every rank keeps a buffer in
// symmetric memory (same layout on every GPU, peers' base pointers known); one single-CTA kernel
// on the compute stream
//   1. publishes "my slot for call #seq is complete" by storing seq into each peer's flag word
//      for this rank (st.release.sys after a system-scope fence),
//   2. spins until every peer has published seq in this rank's flag words (ld.acquire.sys),
//   3. sums the ranks' slots in rank order with system-scope loads — every rank adds the same
//      numbers in the same order, so all ranks hold bitwise identical totals.
// Slots rotate (kPeerSlots): a rank can run at most one call ahead of the slowest peer (it needs
// that peer's flag for the call it is in), so a slot is never overwritten while still being read.
//
// Buffer layout, identical on every rank (all offsets in bytes from the symmetric base):
//   [0, 1024)                     flags: uint32 flag[r] = last call for which rank r's slot is ready
//   1024 + s * slot_bytes         slot s, s in [0, kPeerSlots)
#include "common.cuh"

namespace meb200 {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr uint64_t kPeerTimeoutNs = 30ull * 1000 * 1000 * 1000;   // 30 s

__global__ void __launch_bounds__(256)
k_peer_allreduce_f64(uint8_t *const *__restrict__ bases, uint64_t slot_offset, uint32_t n,
                     uint32_t seq, uint32_t rank, uint32_t world, double *__restrict__ out) {
  const uint32_t tid = threadIdx.x;
  // the slot was filled by earlier kernels of this stream: make it visible system-wide first
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    st_release_sys(reinterpret_cast<uint32_t *>(bases[tid]) + rank, seq);       // tell peer `tid`
    const uint32_t *mine = reinterpret_cast<const uint32_t *>(bases[rank]) + tid;
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) {                         // hear from it
      __nanosleep(100);
      // a peer that died or skipped the call must not hang this GPU for ever: fail loudly
      if (globaltimer_ns() - t0 > kPeerTimeoutNs) __trap();
    }
  }
  __syncthreads();
  for (uint32_t i = tid; i < n; i += blockDim.x) {
    double t = 0.0;
    for (uint32_t r = 0; r < world; ++r)
      t += ld_relaxed_sys_f64(reinterpret_cast<const double *>(bases[r] + slot_offset) + i);
    out[i] = t;
  }
}

// The forward exchange of synchronised batch norm with its consumer fused in: the CTA adds its
// own row count to the slot, exchanges (as above), and turns the totals straight into mean /
// 1/std / running statistics — the work of k_bn_finalize (batchnorm.cu) without a second launch
// and without a separate fill for the count.  slot = [sum x (C) | sum x^2 (C) | rows (1)].
__global__ void __launch_bounds__(256)
k_peer_bn_finalize(uint8_t *const *__restrict__ bases, uint64_t slot_offset, uint32_t C,
                   uint32_t seq, uint32_t rank, uint32_t world, double local_rows, float eps,
                   float momentum, float *__restrict__ running_mean,
                   float *__restrict__ running_var, float *__restrict__ mean,
                   float *__restrict__ invstd, double *__restrict__ total_rows) {
  const uint32_t tid = threadIdx.x;
  __shared__ double s_count;
  if (tid == 0) reinterpret_cast<double *>(bases[rank] + slot_offset)[2 * C] = local_rows;
  __threadfence_system();
  __syncthreads();
  if (tid < world) {
    st_release_sys(reinterpret_cast<uint32_t *>(bases[tid]) + rank, seq);
    const uint32_t *mine = reinterpret_cast<const uint32_t *>(bases[rank]) + tid;
    const uint64_t t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(mine) - seq) < 0) {
      __nanosleep(100);
      if (globaltimer_ns() - t0 > kPeerTimeoutNs) __trap();
    }
  }
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (uint32_t r = 0; r < world; ++r)
      t += ld_relaxed_sys_f64(reinterpret_cast<const double *>(bases[r] + slot_offset) + 2 * C);
    s_count = t;
    *total_rows = t;
  }
  __syncthreads();
  const double count = s_count;
  for (uint32_t c = tid; c < C; c += blockDim.x) {
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t r = 0; r < world; ++r) {
      const double *slot = reinterpret_cast<const double *>(bases[r] + slot_offset);
      s1 += ld_relaxed_sys_f64(slot + c);
      s2 += ld_relaxed_sys_f64(slot + C + c);
    }
    const double mu = s1 / count;
    double var = s2 / count - mu * mu;
    if (var < 0) var = 0;
    mean[c] = (float)mu;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr) {
      const double unbiased = count > 1 ? var * count / (count - 1) : var;
      running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mu);
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
  }
}

}  // namespace meb200

using namespace meb200;

extern "C" {

int meb200_peer_allreduce_f64(const void *peer_bases_dev, uint64_t slot_offset_bytes, uint32_t n,
                              uint32_t seq, uint32_t rank, uint32_t world, double *out,
                              void *stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  MEB_CHECK_ARG(peer_bases_dev != nullptr && out != nullptr, "peer all-reduce: null buffer");
  MEB_CHECK_ARG(world >= 1 && world <= 256 && rank < world, "peer all-reduce: rank %u of %u",
                (unsigned)rank, (unsigned)world);
  MEB_CHECK_ARG(slot_offset_bytes >= 1024 && slot_offset_bytes % 8 == 0 && seq != 0,
                "peer all-reduce: slot offset / sequence number");
  if (n == 0) return MEB200_OK;
  k_peer_allreduce_f64<<<1, 256, 0, s>>>((uint8_t *const *)peer_bases_dev, slot_offset_bytes, n,
                                         seq, rank, world, out);
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

int meb200_peer_bn_finalize(const void *peer_bases_dev, uint64_t slot_offset_bytes, uint32_t C,
                            uint32_t seq, uint32_t rank, uint32_t world, double local_rows,
                            float eps, float momentum, float *running_mean, float *running_var,
                            float *mean, float *invstd, double *total_rows, void *stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  MEB_CHECK_ARG(peer_bases_dev && mean && invstd && total_rows, "peer finalize: null buffer");
  MEB_CHECK_ARG(world >= 1 && world <= 256 && rank < world, "peer finalize: rank %u of %u",
                (unsigned)rank, (unsigned)world);
  MEB_CHECK_ARG(slot_offset_bytes >= 1024 && slot_offset_bytes % 8 == 0 && seq != 0 && C > 0,
                "peer finalize: slot offset / sequence number / channels");
  k_peer_bn_finalize<<<1, 256, 0, s>>>((uint8_t *const *)peer_bases_dev, slot_offset_bytes, C, seq,
                                       rank, world, local_rows, eps, momentum, running_mean,
                                       running_var, mean, invstd, total_rows);
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

}
