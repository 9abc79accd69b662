// Two small HBM-bound kernels around the loss kernels.
//
// row_lse_kernel   scope row f-2 (consume logits): the reference's head ends in log_softmax
//                  (src/models/heads.py:36-37) and its loss starts from those log-probs (losses.py:53-56).
//                  From logits the loss kernels only need ONE number per frame, lse_t = log sum_v exp(x_tv):
//                  p_tv = exp(x_tv - lse_t), and the gradient they already write, exp(lp) - posterior, IS the
//                  gradient with respect to the logits (its rows sum to zero, so log_softmax's backward is
//                  the identity on it).  This kernel reads the logits once (V*4 B/frame) and writes 4 B/frame;
//                  the (B,T,V) log-prob tensor, its gradient and log_softmax's backward pass never exist.
//                  Layout mirrors greedy.cu: one CTA per (utterance, 128-frame tile), tile staged by one 1-D
//                  TMA bulk copy, one thread per frame.
// scale_if_kernel  autograd hands the loss a scalar upstream gradient.  The fused loss launch has already
//                  written d loss / d log_probs with the scalar it EXPECTS (lambda_ctc of losses.py:104, known
//                  on the host before the launch); this kernel reads the actual scalar on the device and
//                  rescales the gradient in place only when it differs -- no host synchronisation, and no
//                  98 MB read+write pass in the common case.
#include <cmath>

#include "ssd_common.cuh"

namespace ssd {

constexpr int LSE_THREADS = 128;  // frames per tile == threads per CTA

__global__ void __launch_bounds__(LSE_THREADS) row_lse_kernel(const float* __restrict__ x, int64_t T, int V, int64_t sb, int64_t st,
                                                              float* __restrict__ lse, int64_t lse_stride, int tiles_per_utt, int use_tma) {
  SSD_DYN_SMEM(smem_raw);
  float* tile = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bar;
  const int tid = (int)threadIdx.x;
  const int64_t b = (int64_t)blockIdx.x / tiles_per_utt;
  const int k = (int)((int64_t)blockIdx.x - b * tiles_per_utt);
  const int64_t t0 = (int64_t)k * LSE_THREADS;
  const int rows = (int)(T - t0 < LSE_THREADS ? T - t0 : LSE_THREADS);
  const float* base = x + b * sb + t0 * st;
  if (use_tma) {
    if (tid == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
      const uint32_t bytes = (uint32_t)(rows * V * 4);
      mbar_arrive_expect_tx(&bar, bytes);
      bulk_g2s(tile, base, bytes, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0u);
  } else {
    for (int i = tid; i < rows * V; i += LSE_THREADS) {
      const int r = i / V, v = i - r * V;
      tile[i] = ld_stream(base + (int64_t)r * st + v);
    }
    __syncthreads();
  }
  if (tid < rows) {
    const float* row = tile + tid * V;
    // each lane starts at a different column so that the 32 rows of a warp (V*4 bytes apart) spread over the banks
    int c = tid % V;
    float m = -INFINITY;
    for (int j = 0; j < V; ++j) {
      m = fmaxf(m, row[c]);
      c = (c + 1 == V) ? 0 : c + 1;
    }
    float out;
    if (!(fabsf(m) <= 3.0e38f)) {
      out = m;  // all -inf, +inf or NaN somewhere: the loss kernels flag the row (|x - lse| is not a small number)
    } else {
      float s = 0.f;
      for (int j = 0; j < V; ++j) {
        s += expf(row[c] - m);
        c = (c + 1 == V) ? 0 : c + 1;
      }
      out = m + logf(s);
    }
    lse[b * lse_stride + t0 + tid] = out;
  }
}

int row_lse_launch(const float* x, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st, float* lse, int64_t lse_stride,
                   cudaStream_t stream) {
  if (B == 0 || T == 0) return SSD_OK;
  const size_t smem = (size_t)LSE_THREADS * V * sizeof(float);
  const int use_tma = (st == V) && aligned16(x) && ((sb * 4) % 16 == 0) && ((V * 4) % 16 == 0) ? 1 : 0;
  const int64_t tiles = (T + LSE_THREADS - 1) / LSE_THREADS;
  if (B * tiles > 2147483647LL) return fail(SSD_ERR_UNSUPPORTED, "row_lse: B*T too large for one grid");
#ifndef SSD_EMU
  cudaError_t e = cudaFuncSetAttribute(row_lse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LSE_THREADS * SSD_MAX_VOCAB * 4);
  if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "row_lse: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
#endif
  SSD_LAUNCH(row_lse_kernel, (unsigned)(B * tiles), LSE_THREADS, smem, stream, x, T, (int)V, sb, st, lse, lse_stride, (int)tiles, use_tma);
  return check_launch("row_lse_kernel");
}

// x[i] *= (*g / expected) for all i < n, unless *g == expected (then nothing is read or written but *g)
__global__ void __launch_bounds__(256) scale_if_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ g, float expected) {
  const float gv = *g;
  if (gv == expected) return;
  const float f = gv / expected;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) {
    float4* x4 = reinterpret_cast<float4*>(x);
    for (int64_t i = i0; i < n4; i += stride) {
      float4 v = x4[i];
      v.x *= f; v.y *= f; v.z *= f; v.w *= f;
      x4[i] = v;
    }
    for (int64_t i = 4 * n4 + i0; i < n; i += stride) x[i] *= f;
  } else {
    for (int64_t i = i0; i < n; i += stride) x[i] *= f;
  }
}

// out[0] = sum_b nll[b] / max(target_lengths[b], 1) in fp64, out[1] = B: the two scalars of the batch-mean loss
// (torch's reduction='mean' of nn.CTCLoss, reference losses.py:28,56) and of its data-parallel all-reduce.  One CTA,
// fixed summation order: the same bits on every run.
__global__ void __launch_bounds__(256) loss_parts_kernel(const float* __restrict__ nll, const int64_t* __restrict__ tg_len, int64_t B,
                                                         double* __restrict__ out) {
  __shared__ double part[256];
  const int tid = (int)threadIdx.x;
  double acc = 0.0;
  for (int64_t b = tid; b < B; b += 256) {
    const int64_t s = tg_len[b];
    acc += (double)nll[b] / (double)(s > 1 ? s : 1);
  }
  part[tid] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) part[tid] += part[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    out[0] = part[0];
    out[1] = (double)B;
  }
}

int loss_parts_launch(const float* nll, const int64_t* tg_len, int64_t B, double* out, cudaStream_t stream) {
  SSD_LAUNCH(loss_parts_kernel, 1u, 256, 0, stream, nll, tg_len, B, out);
  return check_launch("loss_parts_kernel");
}

int scale_if_launch(float* x, int64_t n, const float* g, float expected, cudaStream_t stream) {
  if (n == 0) return SSD_OK;
#ifdef SSD_EMU
  const unsigned grid = 2;
#else
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t want = (n / 4 + 255) / 256;
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > (int64_t)sms * 8 ? (int64_t)sms * 8 : want));
#endif
  SSD_LAUNCH(scale_if_kernel, grid, 256, 0, stream, x, n, g, expected);
  return check_launch("scale_if_kernel");
}

}  // namespace ssd
