// Host-side helpers shared by the C-ABI entry points: error reporting and launch checks.
#pragma once

#include <cstdarg>
#include <cstdio>

#include "../../include/ssd_ctc.h"
#include "ssd_dev.cuh"

namespace ssd {

char* error_buffer();  // thread-local, defined in capi.cu

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  std::vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return SSD_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Entry points implemented one per translation unit.
int greedy_launch(const float* lp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st, const int64_t* lengths, int blank,
                  float bias, int32_t* out_tok, int32_t* out_len, cudaStream_t stream);

size_t beam_workspace_bytes(int64_t B, int64_t T, int64_t V, int width);
size_t beam_lm_workspace_bytes(int64_t B, int64_t T, int64_t V, int width);
// lm == nullptr: the no-LM search
int beam_launch(const float* lp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st, const int64_t* lengths,
                const int32_t* order, int blank, int pad, int space, const ssd_beam_setting* settings, int n_set,
                const ssd_lm_tables* lm, int32_t* out_tok, int32_t* out_len, double* out_score, void* ws, size_t ws_bytes,
                cudaStream_t stream);

size_t loss_workspace_bytes(int64_t B, int64_t T, int64_t V, int64_t S_max);
size_t loss_fallback_counter_offset(int64_t B, int64_t T, int64_t V, int64_t S_max);
int loss_launch(const float* lp, const float* lse, int64_t lse_stride, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st,
                const int64_t* targets, int64_t tgt_stride, int64_t S_max, const int64_t* in_len, const int64_t* tg_len,
                const int32_t* order, int blank, int zero_inf, int64_t mean_den, const float* grad_mul, float* nll, float* grad,
                void* ws, size_t ws_bytes, cudaStream_t stream);

// aux.cu
int row_lse_launch(const float* x, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st, float* lse, int64_t lse_stride,
                   cudaStream_t stream);
int scale_if_launch(float* x, int64_t n, const float* g, float expected, cudaStream_t stream);
int loss_parts_launch(const float* nll, const int64_t* tg_len, int64_t B, double* out, cudaStream_t stream);

int edit_launch(const int32_t* ref_tok, const int64_t* ref_off, const int32_t* hyp_tok, const int64_t* hyp_off, int64_t N,
                int64_t max_ref_len, int64_t max_hyp_len, int32_t* counts, cudaStream_t stream);

}  // namespace ssd
