// SIMT gather-GEMM sparse convolution (output-stationary, no atomics on features).
//
// Reference semantics: ConvolutionForwardKernelCPU / ConvolutionBackwardKernelCPU
// (src/convolution_kernel.hpp:33-144): out[o] = sum_k in[i_k(o)] W_k ;
// dIn[i] = sum_k dOut[o_k(i)] W_k^T ; dW_k = sum_o in[i_k(o)]^T dOut[o].
// Unlike the reference GPU kernels (src/convolution_kernel.cu:114-287: one launch per
// offset, per-element atomicAdd scatter) one launch covers all offsets of a layer and each
// output row is produced by exactly one CTA from the k-major neighbour table.
#include "conv_simt.cuh"

namespace meb200 {

constexpr int TM = 64, TN = 64, TK = 16;

// out[r, n] = sum_k sum_c A[nbr[k][r], c] * Wk(c, n)
//   TRANS_W = false: Wk(c, n) = W[k][c][n]   (W is [K, c_a, c_n])        -> forward
//   TRANS_W = true : Wk(c, n) = W[k][n][c]   (W is [K, c_n, c_a])        -> dgrad
template <typename TIn, typename TOut, bool TRANS_W>
__global__ void __launch_bounds__(256)
k_conv_gather_gemm(const TIn *__restrict__ A, uint32_t c_a, const TIn *__restrict__ W, uint32_t K,
                   uint32_t c_n, const int32_t *__restrict__ nbr, uint32_t n_rows,
                   TOut *__restrict__ out) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  __shared__ int32_t s_idx[TM];

  const uint32_t row0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (uint32_t k = 0; k < K; ++k) {
    __syncthreads();  // previous iteration's readers of s_idx / tiles are done
    int has = 0;
    if (tid < TM) {
      uint32_t r = row0 + tid;
      int32_t v = (r < n_rows) ? __ldg(nbr + (size_t)k * n_rows + r) : -1;
      s_idx[tid] = v;
      has = v >= 0;
    }
    if (!__syncthreads_or(has)) continue;
    const TIn *Wk = W + (size_t)k * c_a * c_n;
    for (uint32_t c0 = 0; c0 < c_a; c0 += TK) {
      // A tile: TM x TK gathered rows (consecutive threads -> consecutive channels)
#pragma unroll
      for (int j = 0; j < (TM * TK) / 256; ++j) {
        int e = tid + j * 256, r = e / TK, c = e % TK;
        int32_t src = s_idx[r];
        float v = 0.f;
        if (src >= 0 && c0 + c < c_a) v = to_f32<TIn>(A[(size_t)src * c_a + c0 + c]);
        As[c][r] = v;
      }
      // B tile: TK x TN of Wk
#pragma unroll
      for (int j = 0; j < (TK * TN) / 256; ++j) {
        int e = tid + j * 256;
        int kk, nn;
        if (TRANS_W) { nn = e / TK; kk = e % TK; } else { kk = e / TN; nn = e % TN; }
        float v = 0.f;
        if (c0 + kk < c_a && n0 + nn < c_n) {
          size_t off = TRANS_W ? (size_t)(n0 + nn) * c_a + (c0 + kk)
                               : (size_t)(c0 + kk) * c_n + (n0 + nn);
          v = to_f32<TIn>(Wk[off]);
        }
        Bs[kk][nn] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t r = row0 + ty * 4 + i;
    if (r >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t n = n0 + tx * 4 + j;
      if (n < c_n) out[(size_t)r * c_n + n] = from_f32<TOut>(acc[i][j]);
    }
  }
}

// dW[k][ci][co] += sum over rows o in this CTA's slice of in[nbr[k][o]][ci] * dOut[o][co]
template <typename T>
__global__ void __launch_bounds__(256)
k_conv_wgrad(const T *__restrict__ in, const T *__restrict__ grad_out, uint32_t c_in,
             uint32_t c_out, const int32_t *__restrict__ out_nbr, uint32_t n_out,
             uint32_t rows_per_split, uint32_t tiles_n, float *__restrict__ dW) {
  __shared__ float As[TK][TM + 4];  // [row in chunk][ci]
  __shared__ float Bs[TK][TN + 4];  // [row in chunk][co]
  __shared__ int32_t s_idx[TK];

  const uint32_t k = blockIdx.x;
  const uint32_t ci0 = (blockIdx.z / tiles_n) * TM, co0 = (blockIdx.z % tiles_n) * TN;
  const uint32_t r_begin = blockIdx.y * rows_per_split;
  const uint32_t r_end = min(r_begin + rows_per_split, n_out);
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (uint32_t r0 = r_begin; r0 < r_end; r0 += TK) {
    __syncthreads();
    int has = 0;
    if (tid < TK) {
      uint32_t r = r0 + tid;
      int32_t v = (r < r_end) ? __ldg(out_nbr + (size_t)k * n_out + r) : -1;
      s_idx[tid] = v;
      has = v >= 0;
    }
    if (!__syncthreads_or(has)) continue;
#pragma unroll
    for (int j = 0; j < (TK * TM) / 256; ++j) {
      int e = tid + j * 256, rr = e / TM, c = e % TM;
      int32_t src = s_idx[rr];
      float a = 0.f, b = 0.f;
      if (src >= 0) {
        if (ci0 + c < c_in) a = to_f32<T>(in[(size_t)src * c_in + ci0 + c]);
        if (co0 + c < c_out) b = to_f32<T>(grad_out[(size_t)(r0 + rr) * c_out + co0 + c]);
      }
      As[rr][c] = a;
      Bs[rr][c] = b;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < TK; ++rr) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[rr][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[rr][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float *dWk = dW + (size_t)k * c_in * c_out;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t ci = ci0 + ty * 4 + i;
    if (ci >= c_in) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t co = co0 + tx * 4 + j;
      if (co < c_out && acc[i][j] != 0.f) atomicAdd(dWk + (size_t)ci * c_out + co, acc[i][j]);
    }
  }
}

// ---- tiny input-channel count (the network stem: c_in <= 4, e.g. RGB -> 32, K = 125) -------
// Too few channels for a GEMM tile; the layer is bound by the neighbour-table scan.  One
// thread per output row keeps all CO outputs in registers; the table column is read
// coalesced, W (fp32, [K, c_in, CO]) is broadcast from shared memory.
template <typename T, typename TOut, int CO>
__global__ void __launch_bounds__(128)
k_conv_small_cin_fwd(const T *__restrict__ in, uint32_t c_in, const T *__restrict__ W, uint32_t K,
                     uint32_t c_out, const int32_t *__restrict__ nbr, uint32_t n_out,
                     TOut *__restrict__ out) {
  extern __shared__ float Ws[];  // [K][c_in][CO], zero padded past c_out
  for (uint32_t e = threadIdx.x; e < K * c_in * CO; e += blockDim.x) {
    uint32_t co = e % CO, kc = e / CO;
    Ws[e] = co < c_out ? to_f32<T>(W[(size_t)kc * c_out + co]) : 0.f;
  }
  __syncthreads();
  const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[CO];
#pragma unroll
  for (int i = 0; i < CO; ++i) acc[i] = 0.f;
  for (uint32_t k = 0; k < K; ++k) {
    int32_t i = (o < n_out) ? __ldg(nbr + (size_t)k * n_out + o) : -1;
    if (i < 0) continue;
    for (uint32_t c = 0; c < c_in; ++c) {
      const float x = to_f32<T>(in[(size_t)i * c_in + c]);
      const float4 *w4 = reinterpret_cast<const float4 *>(Ws + (size_t)(k * c_in + c) * CO);
#pragma unroll
      for (int v = 0; v < CO / 4; ++v) {
        float4 w = w4[v];
        acc[4 * v + 0] = fmaf(x, w.x, acc[4 * v + 0]);
        acc[4 * v + 1] = fmaf(x, w.y, acc[4 * v + 1]);
        acc[4 * v + 2] = fmaf(x, w.z, acc[4 * v + 2]);
        acc[4 * v + 3] = fmaf(x, w.w, acc[4 * v + 3]);
      }
    }
  }
  if (o < n_out) {
#pragma unroll
    for (int i = 0; i < CO; ++i)
      if ((uint32_t)i < c_out) out[(size_t)o * c_out + i] = from_f32<TOut>(acc[i]);
  }
}

// dW[k][c][co] = sum_o in[nbr[k][o]][c] * dOut[o][co] for c_in <= 4, c_out <= 64.
// Warp <-> offset k, lane <-> output channel; a batch of 32 table entries is read coalesced,
// the hits are walked with ballot/shfl; partial sums stay in registers until the row slice ends.
template <typename T>
__global__ void __launch_bounds__(256)
k_conv_small_cin_wgrad(const T *__restrict__ in, const T *__restrict__ gout, uint32_t c_in,
                       uint32_t c_out, uint32_t K, const int32_t *__restrict__ nbr, uint32_t n_out,
                       uint32_t rows_per_block, float *__restrict__ dW) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t k = blockIdx.y * 8 + warp;
  if (k >= K) return;
  const uint32_t r0 = blockIdx.x * rows_per_block;
  const uint32_t r1 = min(r0 + rows_per_block, n_out);
  float acc[4][2] = {};
  const int32_t *nbr_k = nbr + (size_t)k * n_out;
  for (uint32_t o0 = r0; o0 < r1; o0 += 32) {
    const uint32_t o = o0 + lane;
    const int32_t idx = (o < r1) ? __ldg(nbr_k + o) : -1;
    unsigned mask = __ballot_sync(0xffffffffu, idx >= 0);
    while (mask) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      const int32_t i = __shfl_sync(0xffffffffu, idx, b);
      const T *grow = gout + (size_t)(o0 + b) * c_out;
      const float g0 = lane < c_out ? to_f32<T>(grow[lane]) : 0.f;
      const float g1 = lane + 32 < c_out ? to_f32<T>(grow[lane + 32]) : 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if ((uint32_t)c < c_in) {
          const float x = to_f32<T>(in[(size_t)i * c_in + c]);
          acc[c][0] = fmaf(x, g0, acc[c][0]);
          acc[c][1] = fmaf(x, g1, acc[c][1]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if ((uint32_t)c >= c_in) continue;
    float *dst = dW + ((size_t)k * c_in + c) * c_out;
    if (lane < c_out && acc[c][0] != 0.f) atomicAdd(dst + lane, acc[c][0]);
    if (lane + 32 < c_out && acc[c][1] != 0.f) atomicAdd(dst + lane + 32, acc[c][1]);
  }
}

bool conv_small_cin_supported(uint32_t c_in, uint32_t c_out) { return c_in <= 4 && c_out <= 64; }

template <typename T, typename TOut>
static int launch_small_fwd(const void *in, uint32_t c_in, const void *W, uint32_t K,
                            uint32_t c_out, const int32_t *nbr, uint32_t n_out, void *out,
                            cudaStream_t stream) {
  const int CO = c_out <= 32 ? 32 : 64;
  size_t smem = (size_t)K * c_in * CO * sizeof(float);
  if (smem > 200 * 1024) return MEB200_ERR_UNSUPPORTED;
  unsigned grid = cdiv(n_out, 128);
  if (CO == 32) {
    auto kern = k_conv_small_cin_fwd<T, TOut, 32>;
    MEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kern<<<grid, 128, smem, stream>>>((const T *)in, c_in, (const T *)W, K, c_out, nbr, n_out, (TOut *)out);
  } else {
    auto kern = k_conv_small_cin_fwd<T, TOut, 64>;
    MEB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    kern<<<grid, 128, smem, stream>>>((const T *)in, c_in, (const T *)W, K, c_out, nbr, n_out, (TOut *)out);
  }
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

int conv_small_cin_forward(const void *in, int in_dtype, uint32_t c_in, const void *W, uint32_t K,
                           uint32_t c_out, const int32_t *nbr, uint32_t n_out, void *out,
                           int out_dtype, cudaStream_t stream) {
  if (n_out == 0) return MEB200_OK;
  if (in_dtype == MEB200_F32 && out_dtype == MEB200_F32)
    return launch_small_fwd<float, float>(in, c_in, W, K, c_out, nbr, n_out, out, stream);
  if (in_dtype == MEB200_BF16 && out_dtype == MEB200_BF16)
    return launch_small_fwd<__nv_bfloat16, __nv_bfloat16>(in, c_in, W, K, c_out, nbr, n_out, out, stream);
  if (in_dtype == MEB200_BF16 && out_dtype == MEB200_F32)
    return launch_small_fwd<__nv_bfloat16, float>(in, c_in, W, K, c_out, nbr, n_out, out, stream);
  if (in_dtype == MEB200_F16 && out_dtype == MEB200_F16)
    return launch_small_fwd<__half, __half>(in, c_in, W, K, c_out, nbr, n_out, out, stream);
  if (in_dtype == MEB200_F16 && out_dtype == MEB200_F32)
    return launch_small_fwd<__half, float>(in, c_in, W, K, c_out, nbr, n_out, out, stream);
  return MEB200_ERR_UNSUPPORTED;
}

int conv_small_cin_wgrad(const void *in, const void *grad_out, int dtype, uint32_t c_in, uint32_t K,
                         uint32_t c_out, const int32_t *out_nbr, uint32_t n_out, float *grad_weight,
                         cudaStream_t stream) {
  MEB_CUDA(cudaMemsetAsync(grad_weight, 0, (size_t)K * c_in * c_out * sizeof(float), stream));
  if (n_out == 0 || K == 0) return MEB200_OK;
  const uint32_t kgroups = cdiv(K, 8);
  uint32_t want_blocks = cdiv(8ull * num_sms(), kgroups);
  uint32_t rows_per_block = cdiv(cdiv(n_out, want_blocks), 32) * 32;
  if (rows_per_block < 1024) rows_per_block = 1024;
  dim3 grid(cdiv(n_out, rows_per_block), kgroups);
  switch (dtype) {
    case MEB200_F32:
      k_conv_small_cin_wgrad<float><<<grid, 256, 0, stream>>>((const float *)in, (const float *)grad_out, c_in, c_out, K, out_nbr, n_out, rows_per_block, grad_weight);
      break;
    case MEB200_BF16:
      k_conv_small_cin_wgrad<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16 *)in, (const __nv_bfloat16 *)grad_out, c_in, c_out, K, out_nbr, n_out, rows_per_block, grad_weight);
      break;
    case MEB200_F16:
      k_conv_small_cin_wgrad<__half><<<grid, 256, 0, stream>>>((const __half *)in, (const __half *)grad_out, c_in, c_out, K, out_nbr, n_out, rows_per_block, grad_weight);
      break;
    default:
      return MEB200_ERR_UNSUPPORTED;
  }
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

template <typename TIn, typename TOut>
static int launch_gg(const void *A, uint32_t c_a, const void *W, uint32_t K, uint32_t c_n,
                     bool trans_w, const int32_t *nbr, uint32_t n_rows, void *out,
                     cudaStream_t stream) {
  dim3 grid(cdiv(n_rows, TM), cdiv(c_n, TN));
  if (trans_w)
    k_conv_gather_gemm<TIn, TOut, true><<<grid, 256, 0, stream>>>(
        (const TIn *)A, c_a, (const TIn *)W, K, c_n, nbr, n_rows, (TOut *)out);
  else
    k_conv_gather_gemm<TIn, TOut, false><<<grid, 256, 0, stream>>>(
        (const TIn *)A, c_a, (const TIn *)W, K, c_n, nbr, n_rows, (TOut *)out);
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

int conv_forward_simt(const void *in, int in_dtype, uint32_t n_in, uint32_t c_in,
                      const void *weight, uint32_t K, uint32_t c_out, bool trans_w,
                      const int32_t *nbr, uint32_t n_out, void *out, int out_dtype,
                      cudaStream_t stream) {
  (void)n_in;
  if (n_out == 0) return MEB200_OK;
  if (in_dtype == MEB200_F32 && out_dtype == MEB200_F32)
    return launch_gg<float, float>(in, c_in, weight, K, c_out, trans_w, nbr, n_out, out, stream);
  if (in_dtype == MEB200_BF16 && out_dtype == MEB200_F32)
    return launch_gg<__nv_bfloat16, float>(in, c_in, weight, K, c_out, trans_w, nbr, n_out, out, stream);
  if (in_dtype == MEB200_BF16 && out_dtype == MEB200_BF16)
    return launch_gg<__nv_bfloat16, __nv_bfloat16>(in, c_in, weight, K, c_out, trans_w, nbr, n_out, out, stream);
  if (in_dtype == MEB200_F16 && out_dtype == MEB200_F32)
    return launch_gg<__half, float>(in, c_in, weight, K, c_out, trans_w, nbr, n_out, out, stream);
  if (in_dtype == MEB200_F16 && out_dtype == MEB200_F16)
    return launch_gg<__half, __half>(in, c_in, weight, K, c_out, trans_w, nbr, n_out, out, stream);
  set_error("conv (simt): unsupported dtype combination in=%d out=%d", in_dtype, out_dtype);
  return MEB200_ERR_UNSUPPORTED;
}

int conv_wgrad_simt(const void *in, const void *grad_out, int dtype, uint32_t c_in, uint32_t K,
                    uint32_t c_out, const int32_t *out_nbr, uint32_t n_out, float *grad_weight,
                    cudaStream_t stream) {
  MEB_CUDA(cudaMemsetAsync(grad_weight, 0, (size_t)K * c_in * c_out * sizeof(float), stream));
  if (n_out == 0 || K == 0) return MEB200_OK;
  uint32_t tiles_m = cdiv(c_in, TM), tiles_n = cdiv(c_out, TN);
  uint32_t tiles = tiles_m * tiles_n;
  // enough CTAs for ~4 waves; each split covers a multiple of TK rows
  uint32_t want = cdiv(4ull * num_sms(), (uint64_t)K * tiles);
  uint32_t max_splits = cdiv(n_out, 8 * TK);
  uint32_t splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
  if (splits > 65535) splits = 65535;
  uint32_t rows_per_split = cdiv(cdiv(n_out, splits), TK) * TK;
  splits = cdiv(n_out, rows_per_split);
  dim3 grid(K, splits, tiles);
  MEB_CHECK_ARG(tiles <= 65535, "too many channel tiles");
  switch (dtype) {
    case MEB200_F32:
      k_conv_wgrad<float><<<grid, 256, 0, stream>>>((const float *)in, (const float *)grad_out, c_in, c_out, out_nbr, n_out, rows_per_split, tiles_n, grad_weight);
      break;
    case MEB200_BF16:
      k_conv_wgrad<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16 *)in, (const __nv_bfloat16 *)grad_out, c_in, c_out, out_nbr, n_out, rows_per_split, tiles_n, grad_weight);
      break;
    case MEB200_F16:
      k_conv_wgrad<__half><<<grid, 256, 0, stream>>>((const __half *)in, (const __half *)grad_out, c_in, c_out, out_nbr, n_out, rows_per_split, tiles_n, grad_weight);
      break;
    default:
      set_error("conv wgrad (simt): unsupported dtype %d", dtype);
      return MEB200_ERR_UNSUPPORTED;
  }
  MEB_LAUNCH_OK();
  return MEB200_OK;
}

}  // namespace meb200
