// extern "C" entry points of libssd_ctc.so: argument validation, then one kernel launch each.
// Declared in include/ssd_ctc.h.
#include <cmath>

#include "ssd_common.cuh"

namespace ssd {
char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace ssd

using namespace ssd;

static int check_common(const char* fn, const float* lp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st) {
  if (B < 0 || T < 0 || V <= 0) return fail(SSD_ERR_BAD_ARG, "%s: negative or zero dimension (B=%lld T=%lld V=%lld)", fn, (long long)B, (long long)T, (long long)V);
  if (V > SSD_MAX_VOCAB) return fail(SSD_ERR_UNSUPPORTED, "%s: V=%lld exceeds SSD_MAX_VOCAB=%d", fn, (long long)V, SSD_MAX_VOCAB);
  if (B > 0 && T > 0 && !lp) return fail(SSD_ERR_BAD_ARG, "%s: log_probs is NULL", fn);
  if (st < V && T > 1) return fail(SSD_ERR_BAD_ARG, "%s: stride_t=%lld < V=%lld (vocabulary must be contiguous)", fn, (long long)st, (long long)V);
  if (sb < 0 || st < 0) return fail(SSD_ERR_BAD_ARG, "%s: negative stride", fn);
  if (B > 2147483647LL || T > 2147483647LL / (V > 0 ? V : 1)) return fail(SSD_ERR_UNSUPPORTED, "%s: B or T*V exceeds 32-bit range", fn);
  return SSD_OK;
}

extern "C" {

int ssd_abi_version(void) { return SSD_ABI_VERSION; }
const char* ssd_last_error(void) { return error_buffer(); }

int ssd_ctc_greedy(const float* log_probs, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                   const int64_t* lengths, int blank, float blank_bias, int32_t* out_tokens, int32_t* out_lengths,
                   void* stream) {
  int rc = check_common("ssd_ctc_greedy", log_probs, B, T, V, stride_b, stride_t);
  if (rc) return rc;
  if (B > 0 && (!lengths || !out_lengths || (T > 0 && !out_tokens))) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_greedy: NULL lengths/output pointer");
  // blank outside [0,V) is legal for the reference (no id ever equals it); keep that behaviour
  return greedy_launch(log_probs, B, T, V, stride_b, stride_t, lengths, blank, blank_bias, out_tokens, out_lengths, (cudaStream_t)stream);
}

size_t ssd_ctc_beam_workspace_bytes(int64_t B, int64_t T, int64_t V, int beam_width) {
  if (B <= 0 || T < 0 || V <= 0 || beam_width <= 0) return 256;
  return beam_workspace_bytes(B, T, V, beam_width);
}

int ssd_ctc_beam(const float* log_probs, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                 const int64_t* lengths, const int32_t* order, int blank, int pad, int space, float blank_bias,
                 int beam_width, float beam_prune_logp, float token_min_logp, int32_t* out_tokens,
                 int32_t* out_lengths, double* out_scores, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common("ssd_ctc_beam", log_probs, B, T, V, stride_b, stride_t);
  if (rc) return rc;
  if (B > 0 && (!lengths || !out_lengths || (T > 0 && !out_tokens))) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam: NULL lengths/output pointer");
  if (blank < 0 || blank >= V) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam: blank=%d outside [0,%lld)", blank, (long long)V);
  if (beam_width < 1 || beam_width > SSD_MAX_BEAM) return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_beam: beam_width=%d outside [1,%d]", beam_width, SSD_MAX_BEAM);
  if (!(beam_prune_logp <= 0.f)) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam: beam_prune_logp must be <= 0");
  ssd_beam_setting one;
  one.beam_width = beam_width; one.beam_prune_logp = beam_prune_logp; one.blank_bias = blank_bias; one.token_min_logp = token_min_logp;
  return beam_launch(log_probs, B, T, V, stride_b, stride_t, lengths, order, blank, pad, space, &one, 1, nullptr, out_tokens, out_lengths,
                     out_scores, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ssd_ctc_beam_grid(const float* log_probs, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                      const int64_t* lengths, const int32_t* order, int blank, int pad, int space,
                      const ssd_beam_setting* settings, int n_settings, int32_t* out_tokens, int32_t* out_lengths,
                      double* out_scores, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common("ssd_ctc_beam_grid", log_probs, B, T, V, stride_b, stride_t);
  if (rc) return rc;
  if (n_settings < 0 || n_settings > SSD_MAX_BEAM_SETTINGS)
    return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_beam_grid: n_settings=%d outside [0,%d]", n_settings, SSD_MAX_BEAM_SETTINGS);
  if (n_settings > 0 && !settings) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_grid: settings is NULL");
  if (B > 0 && n_settings > 0 && (!lengths || !out_lengths || (T > 0 && !out_tokens))) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_grid: NULL lengths/output pointer");
  if (blank < 0 || blank >= V) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_grid: blank=%d outside [0,%lld)", blank, (long long)V);
  if (B * (int64_t)(n_settings > 0 ? n_settings : 1) > 2147483647LL) return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_beam_grid: B * n_settings exceeds 32-bit range");
  for (int i = 0; i < n_settings; ++i) {
    if (settings[i].beam_width < 1 || settings[i].beam_width > SSD_MAX_BEAM)
      return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_beam_grid: settings[%d].beam_width=%d outside [1,%d]", i, settings[i].beam_width, SSD_MAX_BEAM);
    if (!(settings[i].beam_prune_logp <= 0.f)) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_grid: settings[%d].beam_prune_logp must be <= 0", i);
  }
  return beam_launch(log_probs, B, T, V, stride_b, stride_t, lengths, order, blank, pad, space, settings, n_settings, nullptr, out_tokens,
                     out_lengths, out_scores, workspace, workspace_bytes, (cudaStream_t)stream);
}

uint64_t ssd_lm_hash(uint64_t key) {
  const uint64_t h = key * 0x9E3779B97F4A7C15ull;
  return h ^ (h >> 29);
}

size_t ssd_ctc_beam_lm_workspace_bytes(int64_t B, int64_t T, int64_t V, int beam_width) {
  if (B <= 0 || T < 0 || V <= 0 || beam_width <= 0) return 256;
  return beam_lm_workspace_bytes(B, T, V, beam_width);
}

int ssd_ctc_beam_lm(const float* log_probs, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                    const int64_t* lengths, const int32_t* order, int blank, int pad, int space,
                    const ssd_beam_setting* settings, int n_settings, const ssd_lm_tables* lm, int32_t* out_tokens,
                    int32_t* out_lengths, double* out_scores, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_common("ssd_ctc_beam_lm", log_probs, B, T, V, stride_b, stride_t);
  if (rc) return rc;
  if (n_settings < 0 || n_settings > SSD_MAX_BEAM_SETTINGS)
    return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_beam_lm: n_settings=%d outside [0,%d]", n_settings, SSD_MAX_BEAM_SETTINGS);
  if (n_settings > 0 && !settings) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_lm: settings is NULL");
  if (B > 0 && n_settings > 0 && (!lengths || !out_lengths || (T > 0 && !out_tokens))) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_lm: NULL lengths/output pointer");
  if (blank < 0 || blank >= V) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_lm: blank=%d outside [0,%lld)", blank, (long long)V);
  if (B * (int64_t)(n_settings > 0 ? n_settings : 1) > 2147483647LL) return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_beam_lm: B * n_settings exceeds 32-bit range");
  for (int i = 0; i < n_settings; ++i) {
    if (settings[i].beam_width < 1 || settings[i].beam_width > SSD_MAX_BEAM)
      return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_beam_lm: settings[%d].beam_width=%d outside [1,%d]", i, settings[i].beam_width, SSD_MAX_BEAM);
    if (!(settings[i].beam_prune_logp <= 0.f)) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_lm: settings[%d].beam_prune_logp must be <= 0", i);
  }
  if (!lm) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_lm: lm is NULL (ssd_ctc_beam_grid is the no-LM entry)");
  if (lm->order < 1 || lm->order > SSD_MAX_LM_ORDER)
    return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_beam_lm: LM order %d outside [1,%d]", lm->order, SSD_MAX_LM_ORDER);
  if (!lm->trie_child || !lm->trie_word || !lm->ngram_keys || !lm->ngram_vals || !lm->ngram_prob || !lm->ngram_backoff)
    return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_lm: NULL table pointer in ssd_lm_tables");
  if (lm->n_trie_nodes < 1 || lm->n_ngrams < 1 || (lm->ngram_mask & (lm->ngram_mask + 1)) != 0)
    return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_lm: empty tables or ngram_mask + 1 not a power of two");
  if (lm->bos >= lm->n_ngrams || lm->eos >= lm->n_ngrams || lm->unk >= lm->n_ngrams)
    return fail(SSD_ERR_BAD_ARG, "ssd_ctc_beam_lm: special word id outside the n-gram table");
  return beam_launch(log_probs, B, T, V, stride_b, stride_t, lengths, order, blank, pad, space, settings, n_settings, lm, out_tokens,
                     out_lengths, out_scores, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t ssd_ctc_loss_workspace_bytes(int64_t B, int64_t T, int64_t V, int64_t S_max) {
  if (B <= 0 || T < 0 || V <= 0 || S_max < 0) return 256;
  return loss_workspace_bytes(B, T, V, S_max);
}

size_t ssd_ctc_loss_fallback_counter_offset(int64_t B, int64_t T, int64_t V, int64_t S_max) {
  if (B <= 0 || T < 0 || V <= 0 || S_max < 0) return (size_t)-1;
  return loss_fallback_counter_offset(B, T, V, S_max);
}

int ssd_ctc_loss(const float* log_probs, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                 const int64_t* targets, int64_t tgt_stride, int64_t S_max, const int64_t* input_lengths,
                 const int64_t* target_lengths, const int32_t* order, int blank, int zero_infinity,
                 int64_t mean_denominator, const float* grad_mul, float* nll, float* grad, void* workspace,
                 size_t workspace_bytes, void* stream) {
  int rc = check_common("ssd_ctc_loss", log_probs, B, T, V, stride_b, stride_t);
  if (rc) return rc;
  if (B > 0 && (!input_lengths || !target_lengths || !nll)) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss: NULL lengths/nll pointer");
  if (S_max < 0 || (S_max > 0 && B > 0 && !targets)) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss: bad targets (S_max=%lld)", (long long)S_max);
  if (S_max > SSD_MAX_TARGET) return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_loss: S_max=%lld exceeds SSD_MAX_TARGET=%d", (long long)S_max, SSD_MAX_TARGET);
  if (tgt_stride < S_max) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss: tgt_stride < S_max");
  if (blank < 0 || blank >= V) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss: blank=%d outside [0,%lld)", blank, (long long)V);
  if (mean_denominator < 0) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss: mean_denominator < 0");
  return loss_launch(log_probs, nullptr, 0, B, T, V, stride_b, stride_t, targets, tgt_stride, S_max, input_lengths, target_lengths, order,
                     blank, zero_infinity, mean_denominator, grad_mul, nll, grad, workspace, workspace_bytes, (cudaStream_t)stream);
}

static size_t lse_offset(int64_t B, int64_t T, int64_t V, int64_t S_max) { return (loss_workspace_bytes(B, T, V, S_max) + 255) & ~(size_t)255; }
static int64_t lse_row_stride(int64_t T) { return (T + 7) & ~(int64_t)7; }  // a block of 8 frames = one aligned 32-byte bulk copy

size_t ssd_ctc_loss_from_logits_workspace_bytes(int64_t B, int64_t T, int64_t V, int64_t S_max) {
  if (B <= 0 || T < 0 || V <= 0 || S_max < 0) return 256;
  return lse_offset(B, T, V, S_max) + (size_t)B * (size_t)lse_row_stride(T) * sizeof(float) + 256;
}

int ssd_row_logsumexp(const float* x, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t, float* lse,
                      int64_t lse_stride, void* stream) {
  int rc = check_common("ssd_row_logsumexp", x, B, T, V, stride_b, stride_t);
  if (rc) return rc;
  if (B > 0 && T > 0 && !lse) return fail(SSD_ERR_BAD_ARG, "ssd_row_logsumexp: lse is NULL");
  if (lse_stride < T) return fail(SSD_ERR_BAD_ARG, "ssd_row_logsumexp: lse_stride < T");
  return row_lse_launch(x, B, T, V, stride_b, stride_t, lse, lse_stride, (cudaStream_t)stream);
}

int ssd_ctc_loss_from_logits(const float* logits, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                             const int64_t* targets, int64_t tgt_stride, int64_t S_max, const int64_t* input_lengths,
                             const int64_t* target_lengths, const int32_t* order, int blank, int zero_infinity,
                             int64_t mean_denominator, const float* grad_mul, float* nll, float* grad, void* workspace,
                             size_t workspace_bytes, void* stream) {
  int rc = check_common("ssd_ctc_loss_from_logits", logits, B, T, V, stride_b, stride_t);
  if (rc) return rc;
  if (B > 0 && (!input_lengths || !target_lengths || !nll)) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss_from_logits: NULL lengths/nll pointer");
  if (S_max < 0 || (S_max > 0 && B > 0 && !targets)) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss_from_logits: bad targets (S_max=%lld)", (long long)S_max);
  if (S_max > SSD_MAX_TARGET) return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_loss_from_logits: S_max=%lld exceeds SSD_MAX_TARGET=%d", (long long)S_max, SSD_MAX_TARGET);
  if (tgt_stride < S_max) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss_from_logits: tgt_stride < S_max");
  if (blank < 0 || blank >= V) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss_from_logits: blank=%d outside [0,%lld)", blank, (long long)V);
  if (mean_denominator < 0) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss_from_logits: mean_denominator < 0");
  if (B == 0) return SSD_OK;
  const size_t need = ssd_ctc_loss_from_logits_workspace_bytes(B, T, V, S_max);
  if (!workspace || workspace_bytes < need) return fail(SSD_ERR_WORKSPACE, "ssd_ctc_loss_from_logits: workspace %zu bytes, need %zu", workspace_bytes, need);
  const size_t off = lse_offset(B, T, V, S_max);
  float* lse = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + off);
  const int64_t ls = lse_row_stride(T);
  rc = row_lse_launch(logits, B, T, V, stride_b, stride_t, lse, ls, (cudaStream_t)stream);
  if (rc) return rc;
  return loss_launch(logits, lse, ls, B, T, V, stride_b, stride_t, targets, tgt_stride, S_max, input_lengths, target_lengths, order, blank,
                     zero_infinity, mean_denominator, grad_mul, nll, grad, workspace, off, (cudaStream_t)stream);
}

int ssd_ctc_loss_parts(const float* nll, const int64_t* target_lengths, int64_t B, double* out, void* stream) {
  if (B < 0) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss_parts: B < 0");
  if (!out || (B > 0 && (!nll || !target_lengths))) return fail(SSD_ERR_BAD_ARG, "ssd_ctc_loss_parts: NULL pointer");
  return loss_parts_launch(nll, target_lengths, B, out, (cudaStream_t)stream);
}

int ssd_scale_if(float* x, int64_t n, const float* g, float expected, void* stream) {
  if (n < 0) return fail(SSD_ERR_BAD_ARG, "ssd_scale_if: n < 0");
  if (n > 0 && (!x || !g)) return fail(SSD_ERR_BAD_ARG, "ssd_scale_if: NULL pointer");
  if (!(expected != 0.f) || !(fabsf(expected) <= 3.0e38f)) return fail(SSD_ERR_BAD_ARG, "ssd_scale_if: expected must be finite and non-zero");
  return scale_if_launch(x, n, g, expected, (cudaStream_t)stream);
}

int ssd_edit_counts(const int32_t* ref_tok, const int64_t* ref_off, const int32_t* hyp_tok, const int64_t* hyp_off,
                    int64_t N, int64_t max_ref_len, int64_t max_hyp_len, int32_t* counts, void* stream) {
  if (N < 0) return fail(SSD_ERR_BAD_ARG, "ssd_edit_counts: N=%lld < 0", (long long)N);
  if (N > 0 && (!ref_off || !hyp_off || !counts)) return fail(SSD_ERR_BAD_ARG, "ssd_edit_counts: NULL offsets/counts pointer");
  if (max_ref_len < 0 || max_hyp_len < 0) return fail(SSD_ERR_BAD_ARG, "ssd_edit_counts: negative length bound");
  if (max_ref_len > SSD_MAX_EDIT_LEN || max_hyp_len > SSD_MAX_EDIT_LEN)
    return fail(SSD_ERR_UNSUPPORTED, "ssd_edit_counts: sequences of %lld / %lld tokens exceed SSD_MAX_EDIT_LEN=%d",
                (long long)max_ref_len, (long long)max_hyp_len, SSD_MAX_EDIT_LEN);
  if (N > 0 && ((max_ref_len > 0 && !ref_tok) || (max_hyp_len > 0 && !hyp_tok))) return fail(SSD_ERR_BAD_ARG, "ssd_edit_counts: NULL token pointer");
  if (N > 2147483647LL * 4) return fail(SSD_ERR_UNSUPPORTED, "ssd_edit_counts: N exceeds the grid range");
  return edit_launch(ref_tok, ref_off, hyp_tok, hyp_off, N, max_ref_len, max_hyp_len, counts, (cudaStream_t)stream);
}

}  // extern "C"
