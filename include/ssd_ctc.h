/* ssd_ctc.h -- C ABI of the B200-native CTC hot path (libssd_ctc.so, sm_100a).
 *
 * The reference (ajshul/silent-speech-decoder) is pure Python and has no FFI; the two call
 * sites this library stands behind are
 *
 *   src/models/losses.py:56    self.ctc_loss(log_probs_t, targets, ctc_lengths, target_lengths)
 *                              (nn.CTCLoss(blank=blank_id, zero_infinity=True), :28)  -> ssd_ctc_loss
 *   src/decoding/ctc.py:21-35  torch.argmax + collapse loop (_greedy_decode)           -> ssd_ctc_greedy
 *   src/decoding/ctc.py:75-96  pad->blank merge, label permutation, per-utterance
 *                              pyctcdecode decoder.decode(lp, beam_width, beam_prune_logp) -> ssd_ctc_beam
 *
 * and, for the evaluation loop around the decoder (the first "next" row of the scope table),
 *
 *   src/evaluation/evaluate.py:34-38,61-83  jiwer.wer / jiwer.cer and _levenshtein_counts -> ssd_edit_counts
 *
 * Conventions
 *   - All data pointers are DEVICE pointers (HBM).  The library never allocates, frees or
 *     retains device memory: the caller owns every buffer, including the workspace.
 *   - log_probs is fp32 with element strides (stride_b, stride_t, 1): element (b,t,v) lives at
 *     log_probs[b*stride_b + t*stride_t + v].  The reference's tensors are (B,T,V) contiguous
 *     (heads.py:37) viewed as (T,B,V) (losses.py:53): stride_b = T*V, stride_t = V.
 *   - lengths / ids are int64, as the reference's collate function produces them
 *     (src/data/dataset.py:143,206).
 *   - Work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream) and the call returns without synchronising.
 *   - Every function returns SSD_OK or an error code; ssd_last_error() gives the message of the
 *     last failure on the calling thread.  Nothing throws or aborts across this boundary and
 *     there is no CPU fallback.
 */
#ifndef SSD_CTC_H_
#define SSD_CTC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSD_ABI_VERSION 3

enum {
  SSD_OK = 0,
  SSD_ERR_BAD_ARG = 1,     /* null pointer, negative size, bad stride, id out of range */
  SSD_ERR_UNSUPPORTED = 2, /* shape outside what the kernels are built for */
  SSD_ERR_WORKSPACE = 3,   /* workspace missing or too small */
  SSD_ERR_CUDA = 4         /* a CUDA runtime call or launch failed */
};

/* Limits of this build. */
#define SSD_MAX_VOCAB 128      /* V */
#define SSD_MAX_TARGET 511     /* S: extended length 2S+1 <= 1023 lanes of the alpha/beta recursion */
#define SSD_MAX_BEAM 512       /* beam_width */
#define SSD_MAX_BEAM_SETTINGS 8 /* decoder settings served by one ssd_ctc_beam_grid launch */
#define SSD_MAX_EDIT_LEN 8191  /* tokens per sequence in ssd_edit_counts */
#define SSD_MAX_LM_ORDER 6     /* n-gram order of the language model of ssd_ctc_beam_lm */

int ssd_abi_version(void);
const char* ssd_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Greedy decode: per frame argmax over V (first maximum wins, like torch.argmax on CPU), then
 * collapse: a frame emits its label iff it is not `blank` and differs from the previous frame's
 * argmax (src/decoding/ctc.py:23-35).  blank_bias is added to column `blank` in fp32 before
 * the argmax (ctc.py:18-20).
 *   lengths      [B] frames per utterance, clamped to [0,T]
 *   out_tokens   [B,T] int32, row b holds out_lengths[b] vocabulary ids
 *   out_lengths  [B]   int32
 */
int ssd_ctc_greedy(const float* log_probs, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                   const int64_t* lengths, int blank, float blank_bias, int32_t* out_tokens, int32_t* out_lengths,
                   void* stream);

/* ------------------------------------------------------------------------------------------
 * Prefix beam search with pyctcdecode's no-LM semantics, as invoked by src/decoding/ctc.py:89-96.
 * Done in-kernel per frame: blank column = logaddexp(lp[blank]+blank_bias, lp[pad]) (ctc.py:76-81),
 * the remaining ids in vocabulary order (ctc.py:64-65,84-87), fp32 log-softmax + clip to
 * [ln 1e-15, 0], candidate filter (>= token_min_logp or argmax), fp64 prefix scores, merge by
 * (text, last char), relative prune, top beam_width, history prune on (last word, open word,
 * last char), final merge of blank-/char-ending variants.
 *   pad          vocabulary id folded into blank, or -1
 *   space        vocabulary id of the word separator " ", or -1
 *   order        [B] int32 processing order (e.g. longest first), or NULL for 0..B-1
 *   out_tokens   [B,T] int32 vocabulary ids of the best transcript (whitespace-normalised)
 *   out_lengths  [B]   int32
 *   out_scores   [B]   fp64 log-score of the best transcript, or NULL
 */
size_t ssd_ctc_beam_workspace_bytes(int64_t B, int64_t T, int64_t V, int beam_width);
int ssd_ctc_beam(const float* log_probs, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                 const int64_t* lengths, const int32_t* order, int blank, int pad, int space, float blank_bias,
                 int beam_width, float beam_prune_logp, float token_min_logp, int32_t* out_tokens,
                 int32_t* out_lengths, double* out_scores, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Decoder grid (scope row f-3): the reference sweeps its decoder over (beam_width, beam_prune_logp,
 * blank_bias) points -- src/experiments/config_builder.py:68-93 -- by re-running the whole evaluation
 * once per point (orchestrate.py:129-153,235-249).  Here ONE launch serves up to SSD_MAX_BEAM_SETTINGS
 * points over the same resident batch: a work item is (setting, utterance).  Results are those of
 * n_settings separate ssd_ctc_beam calls, bit for bit.
 *   settings     HOST array [n_settings]
 *   out_tokens   [n_settings, B, T] int32      out_lengths [n_settings, B] int32
 *   out_scores   [n_settings, B] fp64 or NULL
 *   workspace    ssd_ctc_beam_workspace_bytes(B * n_settings, T, V, widest beam_width) bytes
 */
typedef struct ssd_beam_setting {
  int32_t beam_width;
  float beam_prune_logp;
  float blank_bias;
  float token_min_logp;
} ssd_beam_setting;
int ssd_ctc_beam_grid(const float* log_probs, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                      const int64_t* lengths, const int32_t* order, int blank, int pad, int space,
                      const ssd_beam_setting* settings, int n_settings, int32_t* out_tokens, int32_t* out_lengths,
                      double* out_scores, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Prefix beam search fused with a word n-gram language model (scope row f-4): what
 * src/decoding/ctc.py:67-72 sets up when lm_path is given -- pyctcdecode.build_ctcdecoder(labels,
 * kenlm_model_path=<ARPA file of src/decoding/build_char_lm.py:35-40>, alpha, beta) -- and :91-95 runs.
 * A beam is ranked and pruned by  acoustic score + LM score of its closed words + partial-word score:
 *   closed word w after state h   alpha * ln10 * (log10 p(w | h) [+ unk_score_offset if w is out of vocabulary])
 *                                 + beta, back-off scoring as kenlm's BaseScore, start state <s>
 *   open word                     unk_score_offset if it is not a prefix of any unigram, times len/6 beyond 6 chars
 *   end of utterance              the last word also gets log10 p(</s> | ...), unless its text was already
 *                                 scored as a closed text earlier in the utterance (upstream's score cache keyed by
 *                                 text; eos_rescore = 1 selects the other published variant, cache keyed by
 *                                 (text, is_eos): every final beam is re-scored with the </s> term, a text that ends
 *                                 closed scoring an empty -- out-of-vocabulary -- last word)
 *   history prune                 key = (last order-1 closed words, open word, last char)
 * The tables are DEVICE arrays built by the caller from the ARPA file (host side: decoding/lm.py):
 *   trie_child  [n_trie_nodes, n_labels] int32: node reached from `node` by beam label k (label 0 = blank is
 *               unused; labels are numbered as in ssd_ctc_beam: blank, then the vocabulary ids without blank
 *               and pad, ascending), or -1.  Node 0 is the root.  The trie holds the unigrams' spellings.
 *   trie_word   [n_trie_nodes] int32: word id spelled by the path to this node, or -1
 *   ngram_*     forward trie of n-grams as an open-addressing hash table with linear probing: an n-gram
 *               (w1..wn) has an id; unigram ids ARE the word ids; the table maps
 *               key = ((id(w1..w_{n-1}) + 1) << 32) | wn  to id(w1..wn), slot = ssd_lm_hash(key) & ngram_mask,
 *               empty slots hold key 0.  ngram_prob / ngram_backoff [n_ngrams] fp32 log10 values by id.
 *   bos / eos / unk   word ids of <s>, </s>, <unk> (-1 when the file has none)
 * out_scores receives the RANKING score (acoustic + LM) of the best transcript.
 * Limits: order <= SSD_MAX_LM_ORDER; every beam proposes every candidate label (the no-LM kernel's staircase
 * bound does not hold under LM scores), so shared memory limits beam_width to ~100 at 47 labels.
 */
typedef struct ssd_lm_tables {
  const int32_t* trie_child;
  const int32_t* trie_word;
  const uint64_t* ngram_keys;
  const int32_t* ngram_vals;
  const float* ngram_prob;
  const float* ngram_backoff;
  uint64_t ngram_mask; /* slots - 1, slots a power of two */
  int32_t n_trie_nodes, n_labels, n_ngrams, order;
  int32_t bos, eos, unk;
  int32_t eos_rescore; /* 0: a text scored before keeps its score in the final pass (score cache keyed by text);
                          1: the final pass re-scores every beam with the </s> term (cache keyed by (text, is_eos)) */
  double alpha, beta, unk_score_offset;
} ssd_lm_tables;
/* slot hash of the n-gram table: h = key * 0x9E3779B97F4A7C15 (mod 2^64), h ^ (h >> 29).  Exported so that a host
 * builder in any language can check its own copy against the library's. */
uint64_t ssd_lm_hash(uint64_t key);
size_t ssd_ctc_beam_lm_workspace_bytes(int64_t B, int64_t T, int64_t V, int beam_width);
int ssd_ctc_beam_lm(const float* log_probs, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                    const int64_t* lengths, const int32_t* order, int blank, int pad, int space,
                    const ssd_beam_setting* settings, int n_settings, const ssd_lm_tables* lm, int32_t* out_tokens,
                    int32_t* out_lengths, double* out_scores, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * CTC loss, forward and backward in one launch (torch.nn.CTCLoss semantics, reduction 'none'
 * for the returned values; the gradient can be pre-scaled for reduction 'mean').
 *   targets        [B, >=S_max] int64 padded label matrix with row stride tgt_stride (elements)
 *   input_lengths  [B] frames  (<= T)        target_lengths [B] labels (<= S_max)
 *   order          [B] int32 processing order or NULL
 *   nll            [B] fp32: -log p(target | input); +inf if infeasible, 0 if zero_infinity
 *   grad           same layout and strides as log_probs, or NULL for forward only.
 *                  grad[b,t,:] = scale_b * (exp(lp) - posterior), 0 for t >= input_lengths[b] and
 *                  for infeasible rows (zero_infinity).  scale_b = 1 if mean_denominator == 0,
 *                  else 1 / (mean_denominator * max(target_lengths[b],1))  -- pass the GLOBAL
 *                  batch size to obtain d(mean loss)/d(log_probs) without any exchange of
 *                  gradients between data-parallel ranks.
 *   grad_mul       device pointer to one fp32 multiplied into the gradient (autograd's incoming
 *                  grad_output), or NULL for 1
 */
size_t ssd_ctc_loss_workspace_bytes(int64_t B, int64_t T, int64_t V, int64_t S_max);
/* Two kernels serve this entry point: an fp32 fast path that checks its own result per utterance
 * (posterior mass conservation, range flags) and the exact fp64 path, which recomputes whatever
 * the fast path flagged and serves the shapes it does not cover (S_max > 255).  After a call,
 * the uint32 at byte offset ssd_ctc_loss_fallback_counter_offset(...) of the workspace holds how
 * many utterances took the exact path; the offset is (size_t)-1 when the fast path does not
 * apply to the shape (everything is exact then). */
size_t ssd_ctc_loss_fallback_counter_offset(int64_t B, int64_t T, int64_t V, int64_t S_max);
int ssd_ctc_loss(const float* log_probs, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                 const int64_t* targets, int64_t tgt_stride, int64_t S_max, const int64_t* input_lengths,
                 const int64_t* target_lengths, const int32_t* order, int blank, int zero_infinity,
                 int64_t mean_denominator, const float* grad_mul, float* nll, float* grad, void* workspace,
                 size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * The same loss from LOGITS (scope row f-2).  The reference's head ends in log_softmax
 * (src/models/heads.py:36-37) and its loss starts from those log-probs (src/models/losses.py:53-56);
 * this entry consumes the head's pre-softmax output and returns the gradient with respect to it:
 *   grad[b,t,:] = scale_b * (softmax(logits[b,t,:]) - posterior)   (the same numbers ssd_ctc_loss writes:
 *   rows of exp(lp) - posterior sum to zero, so log_softmax's backward is the identity on them).
 * A pre-pass reads the logits once and writes one fp32 log-sum-exp per frame into the workspace; the
 * (B,T,V) log-prob tensor and its gradient never exist.  Arguments as ssd_ctc_loss; the workspace is
 * ssd_ctc_loss_from_logits_workspace_bytes(...) bytes, laid out [ssd_ctc_loss workspace | lse], so
 * ssd_ctc_loss_fallback_counter_offset applies unchanged.
 */
size_t ssd_ctc_loss_from_logits_workspace_bytes(int64_t B, int64_t T, int64_t V, int64_t S_max);
int ssd_ctc_loss_from_logits(const float* logits, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t,
                             const int64_t* targets, int64_t tgt_stride, int64_t S_max, const int64_t* input_lengths,
                             const int64_t* target_lengths, const int32_t* order, int blank, int zero_infinity,
                             int64_t mean_denominator, const float* grad_mul, float* nll, float* grad, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Per-frame log-sum-exp of a (B,T,V) fp32 tensor: lse[b*lse_stride + t] = log sum_v exp(x[b,t,v]).
 * The pre-pass of ssd_ctc_loss_from_logits, exported for callers that keep the values. */
int ssd_row_logsumexp(const float* x, int64_t B, int64_t T, int64_t V, int64_t stride_b, int64_t stride_t, float* lse,
                      int64_t lse_stride, void* stream);

/* The two scalars of the batch-mean loss: out[0] = sum_b nll[b] / max(target_lengths[b], 1) (fp64, fixed summation
 * order), out[1] = B.  mean = out[0] / out[1] is nn.CTCLoss's reduction='mean' (src/models/losses.py:28,56); data-parallel
 * ranks all-reduce the pair.  out: device pointer to two doubles. */
int ssd_ctc_loss_parts(const float* nll, const int64_t* target_lengths, int64_t B, double* out, void* stream);

/* x[i] *= *g / expected for i < n -- unless *g == expected, in which case nothing but *g is read.
 * autograd's backward of the fused loss: the gradient was written with the scalar the caller expects
 * upstream (lambda_ctc of src/models/losses.py:104); the device-side test replaces a host
 * synchronisation, and in the common case the whole read+write pass over the gradient. */
int ssd_scale_if(float* x, int64_t n, const float* g, float expected, void* stream);

/* ------------------------------------------------------------------------------------------
 * Levenshtein alignment counts of N (reference, hypothesis) token sequences: the arithmetic of
 * jiwer.wer / jiwer.cer (src/evaluation/evaluate.py:34-38) and of the in-tree fallback
 * `_levenshtein_counts` (evaluate.py:61-83) -- fewest edits, and among the alignments with the
 * fewest edits the one with the most hits.  Tokens are words (WER) or characters (CER) mapped to
 * non-negative int32 ids by the caller; sequences are concatenated, CSR style.
 *   ref_tok / hyp_tok  int32 ids >= 0          ref_off / hyp_off  [N+1] int64 offsets, off[0] = 0
 *   max_ref_len / max_hyp_len  upper bounds of the sequence lengths (<= SSD_MAX_EDIT_LEN)
 *   counts             [N,4] int32: substitutions, deletions, insertions, hits
 */
int ssd_edit_counts(const int32_t* ref_tok, const int64_t* ref_off, const int32_t* hyp_tok, const int64_t* hyp_off,
                    int64_t N, int64_t max_ref_len, int64_t max_hyp_len, int32_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSD_CTC_H_ */
