// CTC prefix beam search, one CTA per utterance, pyctcdecode (no-LM) semantics.
//
// Replaces reference src/decoding/ctc.py:74-97 (wrapper) + pyctcdecode.BeamSearchDecoderCTC.decode
// (third-party, restated in oracle/beam.py).  See DESIGN.md section 5 for the derivation.
//
// State.  pyctcdecode keys a beam by (text, open word, last char).  For a fixed text the last
// char is either "" (blank) or the last emitted character, so the beam is held as PREFIXES, each
// with two fp64 scores: sb (ends in blank) and snb (ends in its last character) -- but, as
// upstream does, each of the two is ranked, pruned and history-pruned as a separate entry.
// A prefix carries 64-bit rolling hashes of its text (merge key), of its parent's text, of its
// text without a trailing space (final merge), of its last closed word and of its open word
// (history key), its last label, its length and a back-pointer node for the transcript.
// Space handling follows upstream's whitespace-safe merge: a space after "" or after a closed
// word is absorbed (it only changes the last char), so "a" + " " + " " and "a" + " " are one text.
//
// Per frame (4 consumer warps; warp 4 conditions the next frame's row concurrently):
//   1  tot = logaddexp(sb, snb); locate parent prefix through the shared-memory hash table
//   2  rank prefixes by tot; blank / repeat / parent-extension scores of every live prefix
//   3  proposals P+c.  With prefixes ranked by tot (rank i = 0,1,..) and candidate labels by log-prob (column
//      j = 0,1,..), row i keeps its first J_i = floor(W/(i+1)) + 1 columns: about W(ln W + 1.6) proposals instead
//      of W*46.  Proposals whose text is already a live prefix are folded into that prefix's snb (step 2) and
//      dropped here.
//      Why nothing that could reach the top W is dropped.  Take (i,j) with j >= J_i.  Every (i',j') with
//      i' <= i, j' <= j other than (i,j) itself has base tot_{i'} >= tot_i and log-prob r_{j'} >= r_j, so it
//      scores at least as much -- EXCEPT the one column per row i' whose label equals that prefix's last
//      character: that proposal is scored from sb_{i'} (not tot_{i'}), or, for an absorbed space, is not a
//      proposal at all.  A folded proposal does not weaken the count: it lives on as the char-ended entry of
//      its child, whose score lse(rep, ext) >= ext is at least the proposal's, and distinct (i',j') fold into
//      distinct children.  So at least i*j + (j-1) distinct entries score >= (i,j): rows above contribute
//      j+1 columns minus at most one each, row i contributes j columns minus at most one.  With
//      j >= J_i = ceil((W+1)/(i+1)) that is >= (i+1)*j - 1 >= W entries ahead of (i,j): it cannot be among the
//      top W (exact ties aside, which the parity inputs avoid).
//   4  block max, relative prune, exact top-W by rank counting (ties: lower candidate index)
//   5  history prune: best entry per (last closed word, open word, last char)
//   6  compact survivors into the next prefix table, append back-pointer nodes, rebuild hash table
// End: merge blank-/char-ending and trailing-space variants, arg-max, walk back-pointers.
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "ssd_common.cuh"

namespace ssd {

#ifndef SSD_BEAM_CHUNK
#define SSD_BEAM_CHUNK 8
#endif
#ifndef SSD_BEAM_TSMUL
#define SSD_BEAM_TSMUL 4  // hash table slots per beam entry (pow2 above it)
#endif
#ifndef SSD_BEAM_MINCTAS
#define SSD_BEAM_MINCTAS 8
#endif
#ifndef SSD_BEAM_ROWDIV
#define SSD_BEAM_ROWDIV 0  // no-LM kernel, uniform proposal rows: row = index / width instead of the binary search over row offsets
                           // (measured on config 3: 184.1 k against 187.1 k utterances/s confident, flat equal; the LM kernel always divides)
#endif
#ifndef SSD_BEAM_UNROLL1
#define SSD_BEAM_UNROLL1 1
#endif
#if SSD_BEAM_UNROLL1 && !defined(SSD_EMU)
#define SSD_NO_UNROLL _Pragma("unroll 1")
#else
#define SSD_NO_UNROLL
#endif
#ifndef SSD_BEAM_UNROLL1_INNER
#define SSD_BEAM_UNROLL1_INNER 1
#endif
#if SSD_BEAM_UNROLL1_INNER && !defined(SSD_EMU)
#define SSD_NO_UNROLL_INNER _Pragma("unroll 1")
#else
#define SSD_NO_UNROLL_INNER
#endif
// the two rank-counting loops only flat posteriors reach (staircase ranks, boundary bucket of the W-best cut): compute
// loops that gain from the compiler's unrolling (flat config 3: 141 k utterances/s unrolled, 127 k rolled)
#ifndef SSD_BEAM_UNROLL1_FLAT
#define SSD_BEAM_UNROLL1_FLAT 0
#endif
#if SSD_BEAM_UNROLL1_FLAT && !defined(SSD_EMU)
#define SSD_NO_UNROLL_FLAT _Pragma("unroll 1")
#else
#define SSD_NO_UNROLL_FLAT
#endif
#ifndef SSD_BEAM_LM_P3U
#define SSD_BEAM_LM_P3U 1
#endif
#ifndef SSD_BEAM_LSE_INLINE
#define SSD_BEAM_LSE_INLINE 1
#endif
#ifndef SSD_BEAM_CONS
#define SSD_BEAM_CONS 64
#endif
constexpr int BM_CONS = SSD_BEAM_CONS;    // consumer threads (a per-frame phase rarely has more than ~100 items)
constexpr int BM_THREADS = BM_CONS + 32;  // + producer warp
constexpr int BM_CHUNK = SSD_BEAM_CHUNK;              // frames per TMA chunk (two buffers; shared memory decides the CTAs per SM)
#ifndef SSD_BEAM_LM_CHUNK
#define SSD_BEAM_LM_CHUNK 4
#endif
constexpr int BM_LM_CHUNK = SSD_BEAM_LM_CHUNK;        // the LM kernel: smaller chunks, it is after a fourth CTA per SM
constexpr int BAR_CONS = 1;              // named barrier: consumers only
constexpr int BAR_FRAME = 2;             // named barrier: consumers + producer, once per frame
constexpr int LBL_NONE = 255;

struct BeamParams {
  const float* lp;
  int64_t T, sb, st;
  int B, V;
  const int64_t* lengths;
  const int32_t* order;
  int blank, pad, space;
  // decoder settings served by this launch (scope row f-3: the reference re-runs its whole evaluation once per
  // (width, prune, blank bias) point of config_builder.py:68-93; here a work item is (setting, utterance) and all
  // settings read the same resident batch).  W is the CAPACITY = the widest setting: it sizes shared memory.
  int n_set;
  int set_w[SSD_MAX_BEAM_SETTINGS];
  float set_bias[SSD_MAX_BEAM_SETTINGS], set_prune[SSD_MAX_BEAM_SETTINGS], set_tmin[SSD_MAX_BEAM_SETTINGS];
  int W;
  int small_sel;  // selections up to this size take the direct history dedup (tuning knob SSD_BEAM_SMALL)
  int32_t* out_tok;
  int32_t* out_len;
  double* out_score;
  unsigned* counter;     // work queue head
  uint2* nodes;          // [grid][node_cap] back-pointers
  int64_t node_cap;
  int use_tma;
  // derived sizes
  int VL;     // labels after dropping blank/pad: 1 + #others
  int TS;     // hash table slots (pow2)
  int NCAND;  // candidate capacity
  // language-model fusion (beam_kernel<true> only; scope row f-4)
  ssd_lm_tables lm;
  double lm_factor;           // 1 / log10(e): log10 -> ln, as upstream's LOG_BASE_CHANGE_FACTOR
  unsigned long long* seen;   // [grid][seen_cap] closed texts that were LM-scored (upstream's cached_lm_scores keys)
  int64_t seen_cap;           // slots per CTA (pow2)
};

// candidate capacity under LM scores: every live prefix proposes every non-blank label (no staircase bound)
__host__ __device__ inline int beam_ncand_cap_lm(int W, int VL) { return (2 * W + W * (VL - 1) + 31) / 32 * 32; }

__host__ __device__ constexpr int beam_ncand_cap(int W, int VL) {
  // 2 entries per live prefix + staircase sum_{i<W} min(VL-1, W/(i+1)+1)
  long n = 2L * W;
  for (int i = 0; i < W; ++i) {
    int j = W / (i + 1) + 1;
    n += j < VL - 1 ? j : VL - 1;
  }
  return (int)((n + 31) / 32 * 32);
}
__host__ __device__ inline int beam_table_slots(int W) {
  int ts = 64;
  while (ts < SSD_BEAM_TSMUL * W) ts <<= 1;
  return ts;
}

// barrier among the consumer threads: a warp barrier when there is a single consumer warp
template <int CONS>
__device__ __forceinline__ void cons_sync() {
  if (CONS == 32) __syncwarp();
  else bar_sync(BAR_CONS, CONS);
}

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ uint64_t hmix(uint64_t h, uint32_t c) {
  uint64_t x = h ^ (0x9E3779B97F4A7C15ull * (uint64_t)(c + 1u));
  x ^= x >> 32;
  x *= 0xD6E8FEB86659FD93ull;
  x ^= x >> 32;
  x *= 0xD6E8FEB86659FD93ull;
  x ^= x >> 32;
  return x ? x : 1ull;
}
__device__ __forceinline__ double neg_inf() { return -INFINITY; }
// doubles -> unsigned integers in the same order (-inf lowest; no NaN reaches this)
__device__ __forceinline__ unsigned long long order_key(double v) {
  const long long b = __double_as_longlong(v);
  return b >= 0 ? (unsigned long long)b ^ 0x8000000000000000ull : ~(unsigned long long)b;
}
// fp64 logaddexp in upstream's form: hi + log(1 + exp(lo - hi)).  Inlined on purpose: one shared out-of-line copy
// (the fp64 log and exp are ~180 SASS instructions, four copies) cut the registers from 80 to 63 but measured 5 %
// slower at 8 CTAs per SM (185k against 194k utterances/s on config 3, confident input).
#if SSD_BEAM_LSE_INLINE
__device__ __forceinline__ double lse2(double a, double b) {
#else
__device__ __noinline__ double lse2(double a, double b) {
#endif
  const double hi = a >= b ? a : b, lo = a >= b ? b : a;
  if (lo == -INFINITY) return hi;
  return hi + log(1.0 + exp(lo - hi));
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const double x = __shfl_xor_sync(FULL, v, o);
    v = x > v ? x : v;
  }
  return v;
}
// numpy.logaddexp on fp32
__device__ __forceinline__ float logaddexp_f32(float x, float y) {
  if (x == y) return x + 0.69314718055994530942f;
  const float d = x - y;
  if (d > 0.f) return x + log1pf(expf(-d));
  if (d <= 0.f) return y + log1pf(expf(d));
  return x + y;  // NaN
}

struct Prefix {  // structure of arrays, two generations
  uint64_t* H;
  uint64_t* Hpar;
  uint64_t* Hstrip;
  uint64_t* hl;
  uint64_t* ho;
  double* sb;
  double* snb;
  uint32_t* node;
  int* last;  // last label (1..VL-1), or LBL_NONE
  int* len;   // number of labels in the text
};

__device__ __forceinline__ unsigned char* carve(unsigned char*& p, size_t bytes) {
  unsigned char* r = p;
  p += (bytes + 15) & ~(size_t)15;
  return r;
}
__host__ __device__ inline size_t prefix_gen_bytes(int W) {
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  return 7 * al(8 * (size_t)W) + 3 * al(4 * (size_t)W);
}
// The table of generation g starts at base + g * prefix_gen_bytes(W).  An array of two structs indexed by a run-time
// g lives in local memory (20 LDL per thread and frame); pointers rebuilt every frame cost 7 % of the kernel's
// instructions.  So: generation 0's pointers are set up once, and a view adds g's element offset in the subscript.
template <class T>
struct GenArr {
  T* base;
  int off;
  __device__ __forceinline__ T& operator[](int i) const { return base[off + i]; }
};
struct PrefixView {
  GenArr<uint64_t> H, Hpar, Hstrip, hl, ho;
  GenArr<double> sb, snb;
  GenArr<uint32_t> node;
  GenArr<int> last, len;
};
__device__ __forceinline__ PrefixView prefix_view(const Prefix& z, int g, int stride8, int stride4) {
  const int o8 = g * stride8, o4 = g * stride4;
  PrefixView v;
  v.H = {z.H, o8}; v.Hpar = {z.Hpar, o8}; v.Hstrip = {z.Hstrip, o8}; v.hl = {z.hl, o8}; v.ho = {z.ho, o8};
  v.sb = {z.sb, o8}; v.snb = {z.snb, o8};
  v.node = {z.node, o4}; v.last = {z.last, o4}; v.len = {z.len, o4};
  return v;
}
__device__ __forceinline__ Prefix prefix_at(unsigned char* p, int W) {
  Prefix a;
  a.H = (uint64_t*)carve(p, 8 * W);
  a.Hpar = (uint64_t*)carve(p, 8 * W);
  a.Hstrip = (uint64_t*)carve(p, 8 * W);
  a.hl = (uint64_t*)carve(p, 8 * W);
  a.ho = (uint64_t*)carve(p, 8 * W);
  a.sb = (double*)carve(p, 8 * W);
  a.snb = (double*)carve(p, 8 * W);
  a.node = (uint32_t*)carve(p, 4 * W);
  a.last = (int*)carve(p, 4 * W);
  a.len = (int*)carve(p, 4 * W);
  return a;
}

// slots of the per-frame history-key table (pow2 >= 2 * W: load factor at most one half)
__host__ __device__ inline int beam_hk_slots(int W, int NCAND) {
  (void)NCAND;
  int n = 4;
  while (n < 2 * W) n <<= 1;
  return n;
}

// scratch shared by the bucket cut (bucket of each alive entry, boundary list: 2 x NCAND ints) and, after it, the
// history-key table (key, best score, lowest index: 20 bytes per slot)
__host__ __device__ inline size_t beam_scratch_bytes(int W, int NCAND, bool lm = false) {
  // (the LM kernel, whose candidate capacity is 5-10 times larger, keeps the buckets as bytes and the list as 16-bit values)
  const size_t a = (lm ? 3 : 8) * (size_t)NCAND, b = 20 * (size_t)beam_hk_slots(W, NCAND);
  return ((a > b ? a : b) + 15) & ~(size_t)15;
}

// per-prefix language-model state, one generation: accumulated LM score of the closed words (fp64), n-gram context
// ids (order-1, at least one slot), trie node and length of the open word, hashes of the 2nd.. last closed words
__host__ __device__ inline int lm_ctx_slots(int order) { return order > 1 ? order - 1 : 1; }
__host__ __device__ inline size_t lm_gen_bytes(int W, int order) {
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  const int cx = lm_ctx_slots(order);
  return al(8 * (size_t)W) + 2 * al(4 * (size_t)W) + al(4 * (size_t)W * cx) + al(8 * (size_t)W * cx);
}

__host__ __device__ inline size_t beam_smem_bytes(int W, int V, int VL, int TS, int NCAND, int lm_order = 0) {
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  size_t s = 0;
  if (lm_order > 0) s += 2 * lm_gen_bytes(W, lm_order);  // LM state of the prefixes, two generations
  s += 2 * prefix_gen_bytes(W);                              // prefix tables, two generations
  s += 3 * al(8 * (size_t)W);                               // tot, newb, newnb
  s += 5 * al(4 * (size_t)W);                               // par, rank2p, rowoff, keepflags, slot
  s += al(8 * (size_t)TS) + al(4 * (size_t)TS);             // hash table
  s += al(8 * (size_t)NCAND);                               // candidate score (LM: the ranking score)
  if (lm_order == 0) s += al(4 * (size_t)NCAND);            // candidate desc (LM: rows are uniform, desc = f(index))
  s += al((lm_order > 0 ? 2 : 4) * (size_t)NCAND);          // alive idx (LM: 16-bit)
  s += beam_scratch_bytes(W, NCAND, lm_order > 0) + al(4 * 256);  // bucket cut scratch / history-key table, histogram
  s += al(4 * (size_t)(W + 1));                             // the W survivors of the bucket cut
  s += 3 * al(4 * (size_t)(W + 1));                         // sel (cand idx by rank), sel_keep, sel_slot
  s += al(8 * (size_t)(W + 1));                             // sel_hk
  s += 2 * (al(4 * (size_t)VL) + al(VL));                   // frame records: rl[VL], cs[VL] (x2)
  s += al(4 * (size_t)VL);                                  // lab2voc
  s += 2 * al((size_t)(lm_order > 0 ? BM_LM_CHUNK : BM_CHUNK) * V * 4);  // raw log-prob chunks
  s += 256;                                                 // scalars, mbarriers
  return s;
}

struct FrameRec {
  float* rl;          // [VL] conditioned log-prob of label k, -inf when not a candidate
  unsigned char* cs;  // non-blank candidates sorted by log-prob, descending
};

struct Scal {
  uint64_t bar[2];
  double wmax[4];
  int nc1[2];
  int np;
  int n_e3;
  int n_alive;
  int n_b, beta, c0;
  int n_sel;
  int n_keep_old;
  int n_new;
  int utt;
  unsigned nodes_used;
  int wcnt[2][4];  // per-warp counts of the compactions, double-buffered: one barrier per chunk instead of two
};

// ------------------------------------------------------------------ producer: condition one frame
// All 32 lanes; labels k = lane, lane+32, ...  Mirrors ctc.py:76-87 and decode_beams' fp32
// log-softmax + clip, then builds the candidate set {k: r_k >= tmin} | {argmax}.
// PER = label groups of 32 the warp walks (2 for the reference's 47 labels, SSD_MAX_VOCAB / 32 in general); out of line:
// one call per frame, and only the instantiation in use occupies the instruction cache.
#ifndef SSD_BEAM_COND_INLINE
#define SSD_BEAM_COND_INLINE 0
#endif
struct CondArgs {
  int VL, blank, pad;
};
template <int PER>
#if SSD_BEAM_COND_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
void condition_frame_n(const float* __restrict__ row, const CondArgs P, float bias, float tmin,
                                const int* __restrict__ lab2voc, FrameRec fr, int* nc1_out) {
  const int lane = lane_id();
  float x[PER];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int k = lane + 32 * i;
    float v = -INFINITY;
    if (k < P.VL) {
      if (k == 0) {
        v = row[P.blank];
        if (bias != 0.f) v = v + bias;
        if (P.pad >= 0 && P.pad != P.blank) v = logaddexp_f32(v, row[P.pad]);
      } else {
        v = row[lab2voc[k]];
      }
    }
    x[i] = v;
    m = fmaxf(m, v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if (!(fabsf(m) <= 3.0e38f)) m = 0.f;  // non-finite maximum -> 0, as upstream
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int k = lane + 32 * i;
    x[i] = x[i] - m;
    if (k < P.VL) s += (double)expf(x[i]);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  const float lognorm = logf((float)s);
  const float lo = -34.538776394910684f;  // fp32(ln 1e-15)
  float best = -INFINITY;
  int best_k = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int k = lane + 32 * i;
    float r = x[i] - lognorm;
    r = fminf(fmaxf(r, lo), 0.f);
    x[i] = r;
    if (k < P.VL && (r > best || (r == best && k < best_k))) {
      best = r;
      best_k = k;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ob = __shfl_xor_sync(FULL, best, o);
    const int ok = __shfl_xor_sync(FULL, best_k, o);
    if (ob > best || (ob == best && ok < best_k)) {
      best = ob;
      best_k = ok;
    }
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int k = lane + 32 * i;
    if (k < P.VL) fr.rl[k] = (x[i] >= tmin || k == best_k) ? x[i] : -INFINITY;
  }
  __syncwarp();
  // order non-blank candidates by (log-prob desc, label asc): rank = number of candidates ahead.  The loop runs over
  // the CANDIDATES (set bits of a ballot, values handed round by shuffle) -- one or two for a confident frame --
  // not over all VL labels.
  const int ngrp = (P.VL + 31) >> 5;
  bool cand[PER];
  int rank[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int k = lane + 32 * i;
    cand[i] = k >= 1 && k < P.VL && (x[i] >= tmin || k == best_k);
    rank[i] = 0;
  }
  int cnt = 0;
#pragma unroll
  for (int g = 0; g < PER; ++g) {
    if (g >= ngrp) break;  // warp-uniform
    unsigned m = __ballot_sync(FULL, cand[g]);
    cnt += __popc(m);
    while (m) {
      const int b = __ffs((int)m) - 1;
      m &= m - 1u;
      const float o = __shfl_sync(FULL, x[g], b);
      const int q = b + 32 * g;
#pragma unroll
      for (int i = 0; i < PER; ++i) rank[i] += (o > x[i] || (o == x[i] && q < lane + 32 * i)) ? 1 : 0;
    }
  }
#pragma unroll
  for (int i = 0; i < PER; ++i)
    if (cand[i]) fr.cs[rank[i]] = (unsigned char)(lane + 32 * i);
  if (lane == 0) *nc1_out = cnt;
  __syncwarp();
}
__device__ __forceinline__ void condition_frame(const float* __restrict__ row, const BeamParams& P, float bias, float tmin,
                                                const int* __restrict__ lab2voc, FrameRec fr, int* nc1_out) {
  const CondArgs a = {P.VL, P.blank, P.pad};
#if SSD_BEAM_COND_INLINE
  condition_frame_n<SSD_MAX_VOCAB / 32>(row, a, bias, tmin, lab2voc, fr, nc1_out);
#else
  if (P.VL <= 64) condition_frame_n<2>(row, a, bias, tmin, lab2voc, fr, nc1_out);
  else condition_frame_n<SSD_MAX_VOCAB / 32>(row, a, bias, tmin, lab2voc, fr, nc1_out);
#endif
}

// ------------------------------------------------------------------ hash table (shared memory)
__device__ __forceinline__ void table_insert(unsigned long long* keys, int* vals, int mask, uint64_t h, int v) {
  int s = (int)(h & (uint64_t)mask);
  for (;;) {
    const unsigned long long old = atomicCAS(&keys[s], 0ull, (unsigned long long)h);
    if (old == 0ull || old == (unsigned long long)h) {
      vals[s] = v;
      return;
    }
    s = (s + 1) & mask;
  }
}
__device__ __forceinline__ int table_find(const unsigned long long* keys, const int* vals, int mask, uint64_t h) {
  if (h == 0ull) return -1;  // 0 marks an empty slot and "no parent"
  int s = (int)(h & (uint64_t)mask);
  for (;;) {
    const unsigned long long k = keys[s];
    if (k == (unsigned long long)h) return vals[s];
    if (k == 0ull) return -1;
    s = (s + 1) & mask;
  }
}

// rows that a bulk copy cannot take (non-contiguous or unaligned input): a plain coalesced copy by the producer warp.
// Out of line: it is cold code that was instantiated three times inside the hot kernel.
__device__ __noinline__ void load_rows_strided(float* __restrict__ dst, const float* __restrict__ src, int rows, int V, int64_t st) {
  const int lane = lane_id();
  for (int i = lane; i < rows * V; i += 32) {
    const int r = i / V, v = i - r * V;
    dst[i] = ld_stream(src + (int64_t)r * st + v);
  }
  __syncwarp();
}

// ------------------------------------------------------------------ language model (beam_kernel<true>)
// Restates what pyctcdecode's LanguageModel does with a kenlm.Model (oracle/lm.py has the same arithmetic in numpy):
// back-off scoring over a forward trie of n-grams held as a hash table in global memory, unigram spellings as a
// dense character trie.  All tables are read-only and L2-resident for LMs of the reference's size.
constexpr int LM_CX = SSD_MAX_LM_ORDER - 1;  // context words / history words kept per prefix, at most

__device__ __forceinline__ uint64_t lm_hash(uint64_t key) {
  const uint64_t h = key * 0x9E3779B97F4A7C15ull;
  return h ^ (h >> 29);
}
// id of the n-gram (context n-gram `parent`, word w), or -1
__device__ inline int lm_child(const ssd_lm_tables& L, int parent, int w) {
  const uint64_t key = ((uint64_t)(uint32_t)(parent + 1) << 32) | (uint64_t)(uint32_t)w;
  uint64_t s = lm_hash(key) & L.ngram_mask;
  for (uint64_t tries = 0; tries <= L.ngram_mask; ++tries) {  // bounded: a caller's table without an empty slot must not hang the GPU
    const uint64_t k = L.ngram_keys[s];
    if (k == key) return L.ngram_vals[s];
    if (k == 0ull) return -1;
    s = (s + 1) & L.ngram_mask;
  }
  return -1;
}
// kenlm BaseScore: log10 p(w | ctx) in fp32.  ctx[j] = id of the n-gram made of the last j+1 context words (-1: not
// in the model).  The longest n-gram ending in w gives the probability; every context longer than the matched one
// adds its back-off weight, shorter contexts first.  new_ctx (may be null) receives the state after w.
__device__ inline float lm_base_score(const ssd_lm_tables& L, const int* ctx, int w, int* new_ctx) {
  const int C = L.order - 1;
  const int wid = w >= 0 ? w : L.unk;  // out of vocabulary -> <unk>
  float prob = wid >= 0 ? L.ngram_prob[wid] : -100.f;
  int matched = 1;
  if (new_ctx) {
    for (int j = 0; j < C; ++j) new_ctx[j] = -1;
    if (C > 0) new_ctx[0] = wid;
  }
  if (wid >= 0) {
    for (int j = 0; j < C; ++j) {
      if (ctx[j] < 0) break;
      const int id = lm_child(L, ctx[j], wid);
      if (id < 0) break;
      prob = L.ngram_prob[id];
      matched = j + 2;
      if (new_ctx && j + 1 < C) new_ctx[j + 1] = id;
    }
  }
  for (int j = matched - 1; j < C; ++j)
    if (ctx[j] >= 0) prob = prob + L.ngram_backoff[ctx[j]];
  return prob;
}
// LanguageModel.score: alpha * (BaseScore [+ unk offset] [+ BaseScore(</s>)]) * ln10 + beta, in fp64 over fp32 terms
#ifndef SSD_BEAM_LM_WORD_INLINE
#define SSD_BEAM_LM_WORD_INLINE 1
#endif
#if SSD_BEAM_LM_WORD_INLINE
__device__ inline
#else
__device__ __noinline__  // one copy, off the hot path: only space proposals and the end of an utterance score words
#endif
double lm_word_score(const ssd_lm_tables& L, double factor, const int* ctx, int w, bool eos, int* new_ctx) {
  int st[LM_CX];
  double lm = (double)lm_base_score(L, ctx, w, st);
  if (w < 0) lm += L.unk_score_offset;
  if (eos) lm = lm + (double)lm_base_score(L, st, L.eos, nullptr);
  if (new_ctx)
    for (int j = 0; j < L.order - 1; ++j) new_ctx[j] = st[j];
  return L.alpha * lm * factor + L.beta;
}
// score_partial_token: an open word that is no prefix of any unigram costs the unk offset, growing beyond 6 characters
__device__ __forceinline__ double lm_partial_score(const ssd_lm_tables& L, int node, int plen) {
  if (plen == 0) return 0.0;
  double u = node < 0 ? L.unk_score_offset : 0.0;
  if (plen > 6) u = u * (double)plen / 6.0;
  return u;
}
// set of 64-bit keys in global memory (one per CTA): the closed texts that were scored
__device__ inline void seen_insert(unsigned long long* tab, unsigned long long mask, uint64_t h) {
  unsigned long long s = h & mask;
  for (unsigned long long tries = 0; tries <= mask; ++tries) {  // the set is sized for twice what an utterance can insert
    const unsigned long long old = atomicCAS(&tab[s], 0ull, (unsigned long long)h);
    if (old == 0ull || old == (unsigned long long)h) return;
    s = (s + 1) & mask;
  }
}
__device__ inline bool seen_find(unsigned long long* tab, unsigned long long mask, uint64_t h) {
  unsigned long long s = h & mask;
  for (unsigned long long tries = 0; tries <= mask; ++tries) {
    const unsigned long long k = atomicCAS(&tab[s], 0ull, 0ull);  // an L2 read
    if (k == (unsigned long long)h) return true;
    if (k == 0ull) return false;
    s = (s + 1) & mask;
  }
  return false;
}

// ------------------------------------------------------------------ the kernel
// LM = false: the no-LM search (the reference's configured path).  LM = true: the same search ranked by acoustic +
// LM score; every `if constexpr (LM)` block below is absent from the no-LM instantiation.
#ifndef SSD_BEAM_LM_MINCTAS
#define SSD_BEAM_LM_MINCTAS 3  // 164 registers; shared memory (54.5 KB at width 50) then admits a fourth CTA, which the register file still holds
#endif
#ifndef SSD_BEAM_LM_CONS
#define SSD_BEAM_LM_CONS 64  // consumer threads of the LM kernel (a frame has 5-10 times the candidates of a no-LM one)
#endif
constexpr int BM_LM_CONS = SSD_BEAM_LM_CONS;
constexpr int BM_LM_THREADS = BM_LM_CONS + 32;
static_assert(BM_LM_CONS % 32 == 0 && BM_LM_CONS <= 128 && BM_CONS <= 128, "per-warp scratch in Scal holds four consumer warps");
// WCAP: the capacity (widest beam of the launch) as a compile-time constant for the widths the reference uses, 0 = taken
// from the launch.  VLC: likewise the label count (47 = the reference's 48-token vocabulary without pad), which fixes the
// candidate capacity and with it every remaining offset but the raw chunks' (config 3: another +4.7 % / +2.3 %).  With a constant capacity every W-sized table below sits at a constant shared-memory offset; with a
// run-time one the compiler re-derives those addresses inside the frame loop (15 % of the executed instructions).
template <bool LM, int WCAP, int VLC = 0>
__global__ void __launch_bounds__(LM ? BM_LM_THREADS : BM_THREADS, LM ? SSD_BEAM_LM_MINCTAS : SSD_BEAM_MINCTAS) beam_kernel(const BeamParams P) {
  constexpr int CONS = LM ? BM_LM_CONS : BM_CONS;  // consumer threads; one more warp produces
  constexpr int THREADS = CONS + 32;
  SSD_DYN_SMEM(smem_raw);
  unsigned char* sp = smem_raw;
  constexpr int NCAND_C = (VLC > 0 && WCAP > 0 && !LM) ? beam_ncand_cap(WCAP > 0 ? WCAP : 1, VLC > 0 ? VLC : 2) : 0;
  const int VL = VLC > 0 ? VLC : P.VL, NCAND = NCAND_C > 0 ? NCAND_C : P.NCAND;

  const int WC = WCAP > 0 ? WCAP : P.W;  // capacity: every table below is sized for the widest setting of the launch
  const int TS = WCAP > 0 ? beam_table_slots(WCAP) : P.TS;
  unsigned char* const gen_base = sp;
  const size_t gen_stride = prefix_gen_bytes(WC);  // a multiple of 16
  const Prefix gen0 = prefix_at(gen_base, WC);
  const int gs8 = (int)(gen_stride / 8), gs4 = (int)(gen_stride / 4);
  sp += 2 * gen_stride;
  double* tot = (double*)carve(sp, 8 * WC);
  double* newb = (double*)carve(sp, 8 * WC);
  double* newnb = (double*)carve(sp, 8 * WC);
  int* par = (int*)carve(sp, 4 * WC);
  int* rank2p = (int*)carve(sp, 4 * WC);
  int* rowoff = (int*)carve(sp, 4 * WC);
  int* keepf = (int*)carve(sp, 4 * WC);
  int* oslot = (int*)carve(sp, 4 * WC);
  unsigned long long* tkeys = (unsigned long long*)carve(sp, 8 * TS);
  int* tvals = (int*)carve(sp, 4 * TS);
  int* hist = (int*)carve(sp, 4 * 256);
  int* bidx = (int*)carve(sp, 4 * (WC + 1));
  int* sel = (int*)carve(sp, 4 * (WC + 1));
  int* sel_keep = (int*)carve(sp, 4 * (WC + 1));
  int* sel_slot = (int*)carve(sp, 4 * (WC + 1));
  uint64_t* sel_hk = (uint64_t*)carve(sp, 8 * (WC + 1));
  Scal* sc = (Scal*)carve(sp, sizeof(Scal));
  // language-model state of the prefixes (two generations) and the candidates' ranking scores (acoustic + LM)
  const int CX = lm_ctx_slots(P.lm.order);
  double* lmt_base = nullptr;   // accumulated LM score of the closed words
  int* ltrie_base = nullptr;    // trie node of the open word (-1: no unigram starts like it)
  int* lplen_base = nullptr;    // characters in the open word
  int* lctx_base = nullptr;     // [W][CX] n-gram ids of the last 1.. order-1 closed words
  uint64_t* lhw_base = nullptr; // [W][CX] hashes of the 2nd, 3rd, .. last closed words (the last one is Prefix::hl)
  double* cfused = nullptr;
  int lm_gs8 = 0, lm_gs4 = 0;
  if constexpr (LM) {
    unsigned char* const lp0 = sp;
    lmt_base = (double*)carve(sp, 8 * WC);
    ltrie_base = (int*)carve(sp, 4 * WC);
    lplen_base = (int*)carve(sp, 4 * WC);
    lctx_base = (int*)carve(sp, 4 * WC * CX);
    lhw_base = (uint64_t*)carve(sp, 8 * WC * CX);
    const size_t gb = lm_gen_bytes(WC, P.lm.order);
    sp = lp0 + 2 * gb;
    lm_gs8 = (int)(gb / 8);
    lm_gs4 = (int)(gb / 4);
  }
  // ---- from here on the offsets depend on the label count (run-time): candidate arrays, frame records, raw chunks
  double* cscore = (double*)carve(sp, 8 * NCAND);  // LM: the RANKING score (acoustic + LM); the acoustic one is recomputed for survivors
  uint32_t* cdesc = LM ? nullptr : (uint32_t*)carve(sp, 4 * NCAND);
  using ai_t = typename std::conditional<LM, unsigned short, int>::type;  // candidate / alive indices (LM: capacity < 65536)
  using bk_t = typename std::conditional<LM, unsigned char, int>::type;   // bucket of an alive entry
  ai_t* aidx = (ai_t*)carve(sp, sizeof(ai_t) * NCAND);
  bk_t* abkt = (bk_t*)carve(sp, beam_scratch_bytes(WC, NCAND, LM));  // after the cut this scratch holds the history-key table
  ai_t* blist = reinterpret_cast<ai_t*>(abkt + NCAND);
  // bucket codes: no-LM int: bucket b, -1 = in, 256 = out.  LM byte: b + 1 with b <= 253, 0 = in, 255 = out
  constexpr int BK_OFF = LM ? 1 : 0, BK_TOP = LM ? 253 : 255, BK_IN = LM ? 0 : -1, BK_OUT = LM ? 255 : 256;
  constexpr int CHUNK = LM ? BM_LM_CHUNK : BM_CHUNK;
  float* const rl_base = (float*)carve(sp, 4 * VL);
  unsigned char* const cs_base = (unsigned char*)carve(sp, VL);
  float* const rl_1 = (float*)carve(sp, 4 * VL);
  unsigned char* const cs_1 = (unsigned char*)carve(sp, VL);
  const ptrdiff_t rl_stride = rl_1 - rl_base, cs_stride = cs_1 - cs_base;
  auto frame_rec = [&](int i) {
    FrameRec f;
    f.rl = rl_base + (i & 1) * rl_stride;
    f.cs = cs_base + (i & 1) * cs_stride;
    return f;
  };
  int* lab2voc = (int*)carve(sp, 4 * VL);
  float* const chunk0 = (float*)carve(sp, (size_t)CHUNK * P.V * 4);
  float* const chunk1 = (float*)carve(sp, (size_t)CHUNK * P.V * 4);
  const ptrdiff_t chunk_stride = chunk1 - chunk0;
  auto chunk_buf = [&](int c) { return chunk0 + (c & 1) * chunk_stride; };
  // history-key table over the (by then free) abkt/blist scratch: keys, then the best rank per key
  const int HKS = beam_hk_slots(WC, NCAND);
  unsigned long long* hk_keys = reinterpret_cast<unsigned long long*>(abkt);
  unsigned long long* hk_best = hk_keys + HKS;
  int* hk_idx = reinterpret_cast<int*>(hk_best + HKS);
  if constexpr (LM) cfused = cscore;
  const double* const rscore = cscore;  // what ranks and prunes a candidate
  unsigned long long* const seen = LM ? P.seen + (int64_t)blockIdx.x * P.seen_cap : nullptr;
  (void)CX; (void)lm_gs4;

  const int tid = (int)threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const bool producer = warp == CONS / 32;
  const int tmask = TS - 1;
  uint2* nodes = P.nodes + (int64_t)blockIdx.x * P.node_cap;
  const int space_lbl_vocab = P.space;

  // label <-> vocabulary id (ctc.py:64-65): label 0 = blank, then ids in order without blank/pad
  if (tid == 0) {
    int k = 1;
    lab2voc[0] = P.blank;
    for (int v = 0; v < P.V; ++v)
      if (v != P.blank && v != P.pad) lab2voc[k++] = v;
    mbar_init(&sc->bar[0], 1);
    mbar_init(&sc->bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  int space_lbl = -1;
  SSD_NO_UNROLL_INNER
  for (int k = 1; k < VL; ++k)
    if (lab2voc[k] == space_lbl_vocab) space_lbl = k;

  unsigned tma_phase[2] = {0u, 0u};  // producer-side parity per chunk buffer

  for (;;) {
    // ---------------- fetch next utterance
    if (tid == 0) sc->utt = (int)atomicAdd(P.counter, 1u);
    __syncthreads();
    const int u = sc->utt;
    if (u >= P.B * P.n_set) break;
    // work item = (setting, utterance); settings come widest first, utterances longest first within a setting
    const int si = u / P.B, ub = u - si * P.B;
    const int W = P.set_w[si];
    const float s_bias = P.set_bias[si], s_prune = P.set_prune[si], s_tmin = P.set_tmin[si];
    const int b = P.order ? P.order[ub] : ub;
    const int64_t ob = (int64_t)si * P.B + b;  // output row
    int64_t len64 = P.lengths[b];
    const int len = (int)(len64 < 0 ? 0 : (len64 > P.T ? P.T : len64));
    const float* base = P.lp + (int64_t)b * P.sb;
    const int nchunks = (len + CHUNK - 1) / CHUNK;

    // ---------------- initial beam: the empty prefix, blank-ended with score 0
    if (tid == 0) {
      const Prefix& a = gen0;
      a.H[0] = hmix(0x243F6A8885A308D3ull, 0u);
      a.Hpar[0] = 0ull;  // never a key
      a.Hstrip[0] = a.H[0];
      a.hl[0] = 0ull;
      a.ho[0] = 0ull;
      a.sb[0] = 0.0;
      a.snb[0] = neg_inf();
      a.node[0] = 0u;
      a.last[0] = space_lbl >= 0 ? space_lbl : LBL_NONE;  // a space after "" is absorbed
      a.len[0] = 0;
      sc->np = 1;
      sc->nodes_used = 1u;
      nodes[0] = make_uint2(0u, 0u);
      if constexpr (LM) {  // cached_lm_scores = {"": (0, start state <s>)}
        lmt_base[0] = 0.0;
        ltrie_base[0] = 0;
        lplen_base[0] = 0;
        for (int j = 0; j < CX; ++j) {
          lctx_base[j] = (j == 0) ? P.lm.bos : -1;
          lhw_base[j] = 0ull;
        }
      }
    }
    unsigned long long seen_mask = 0ull;
    if constexpr (LM) {
      // at most one new closed text per live prefix and frame; load factor <= 1/2
      unsigned long long slots = 1024ull;
      while (slots < 2ull * (unsigned long long)len * (unsigned long long)W && (int64_t)slots < P.seen_cap) slots <<= 1;
      seen_mask = slots - 1ull;
      SSD_NO_UNROLL
      for (unsigned long long i = (unsigned long long)tid; i < slots; i += THREADS) seen[i] = 0ull;
    }
    SSD_NO_UNROLL
    for (int i = tid; i < TS; i += THREADS) tkeys[i] = 0ull;
    SSD_NO_UNROLL
    for (int i = tid; i < HKS; i += THREADS) {
      hk_keys[i] = 0ull;
      hk_best[i] = 0ull;
      hk_idx[i] = 0x7fffffff;
    }
    bool hk_dirty = false;
    __syncthreads();
    if (tid == 0) table_insert(tkeys, tvals, tmask, gen0.H[0], 0);

    // ---------------- producer prologue: chunk 0 (+1 in flight), frame 0 record
    auto load_chunk = [&](int c) {  // producer warp
      const int rows = min(CHUNK, len - c * CHUNK);
      float* dst = chunk_buf(c);
      const float* src = base + (int64_t)c * CHUNK * P.st;
      if (P.use_tma) {
        if (lane == 0) {
          const uint32_t bytes = (uint32_t)(rows * P.V * 4);
          mbar_arrive_expect_tx(&sc->bar[c & 1], bytes);
          bulk_g2s(dst, src, bytes, &sc->bar[c & 1]);
        }
      } else {
        load_rows_strided(dst, src, rows, P.V, P.st);
      }
    };
    auto wait_chunk = [&](int c) {
      if (P.use_tma) {
        mbar_wait(&sc->bar[c & 1], tma_phase[c & 1]);
        tma_phase[c & 1] ^= 1u;
      }
    };
    if (producer && len > 0) {
      load_chunk(0);
      if (nchunks > 1) load_chunk(1);
      wait_chunk(0);
      condition_frame(chunk0, P, s_bias, s_tmin, lab2voc, frame_rec(0), &sc->nc1[0]);
    }
    __syncthreads();

    int g = 0;  // current prefix generation
    bool hk_clean = false;  // the live prefixes are known to have pairwise different history keys (uniform over the consumers)
#ifdef SSD_BEAM_PROF
    long long pf_prod = 0, pf_fast = 0, pf_gen = 0, pf_wait = 0;
    int pf_nfast = 0, pf_ngen = 0;
    long long pf_ph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long pf_last = 0;
#define PF_STAMP(k) do { const long long pf_now = clock64(); pf_ph[k] += pf_now - pf_last; pf_last = pf_now; } while (0)
#else
#define PF_STAMP(k) do { } while (0)
#endif
    for (int t = 0; t < len; ++t) {
#ifdef SSD_BEAM_PROF
      const long long pf_t0 = clock64();
#endif
      if (producer) {
        // -------- condition frame t+1 while the consumers work on frame t
        const int tn = t + 1;
        if (tn < len) {
          const int c = tn / CHUNK, r = tn - c * CHUNK;
          if (r == 0) {
            // chunk c-1 is fully consumed (its last frame was conditioned an iteration ago)
            if (c + 1 < nchunks) load_chunk(c + 1);
            wait_chunk(c);
          }
          condition_frame(chunk_buf(c) + r * P.V, P, s_bias, s_tmin, lab2voc, frame_rec(tn), &sc->nc1[tn & 1]);
        }
#ifdef SSD_BEAM_PROF
        const long long pf_t1 = clock64();
        pf_prod += pf_t1 - pf_t0;
#endif
        bar_sync(BAR_FRAME, THREADS);
#ifdef SSD_BEAM_PROF
        pf_wait += clock64() - pf_t1;
        if (t == len - 1 && blockIdx.x == 0 && lane == 0)
          printf("producer: %d frames, busy %lld cycles/frame, waiting %lld cycles/frame\n", len, pf_prod / len, pf_wait / len);
#endif
        continue;
      }
      // ============================ consumers ============================
      const PrefixView A = prefix_view(gen0, g, gs8, gs4);
      const PrefixView N = prefix_view(gen0, g ^ 1, gs8, gs4);
      const FrameRec F = frame_rec(t);
      const int nc1 = sc->nc1[t & 1];
      const int np = sc->np;
      // LM state of the current (A) and next (N) generation
      double* const lmtA = LM ? lmt_base + g * lm_gs8 : nullptr;
      double* const lmtN = LM ? lmt_base + (g ^ 1) * lm_gs8 : nullptr;
      int* const trieA = LM ? ltrie_base + g * lm_gs4 : nullptr;
      int* const trieN = LM ? ltrie_base + (g ^ 1) * lm_gs4 : nullptr;
      int* const plenA = LM ? lplen_base + g * lm_gs4 : nullptr;
      int* const plenN = LM ? lplen_base + (g ^ 1) * lm_gs4 : nullptr;
      int* const ctxA = LM ? lctx_base + g * lm_gs4 : nullptr;
      int* const ctxN = LM ? lctx_base + (g ^ 1) * lm_gs4 : nullptr;
      uint64_t* const hwA = LM ? lhw_base + g * lm_gs8 : nullptr;
      uint64_t* const hwN = LM ? lhw_base + (g ^ 1) * lm_gs8 : nullptr;
      // LM score of a live prefix: closed words + open word
      auto lm_of = [&](int p) -> double { return lmtA[p] + lm_partial_score(P.lm, trieA[p], plenA[p]); };
      // history key: (last closed word(s), open word, last char); with an LM the last order-1 closed words count
      // (`older` = hashes of the 2nd, 3rd, .. last closed word, `shifted`: the entry closes a word, so the prefix's
      // last word `hl_old` becomes the 2nd last)
      auto hist_key = [&](uint64_t hl, uint64_t ho, int lc, int p, bool shifted, uint64_t hl_old) -> uint64_t {
        uint64_t hk = hmix(hmix(hl ^ hmix(ho, 0x5bu), 0xa7u), 0x100u + (uint32_t)lc);  // never 0 (hmix)
        if constexpr (LM) {
          for (int j = 0; j + 2 <= CX; ++j) {  // the (j+2)-th last closed word
            const uint64_t w = shifted ? (j == 0 ? hl_old : hwA[p * CX + j - 1]) : hwA[p * CX + j];
            hk = hmix(hk ^ w, 0x33u + (uint32_t)j);
          }
        }
        return hk;
      };
      const unsigned node0 = sc->nodes_used;  // read here: thread 0 rewrites it in phase 6b
      SSD_ASSERT(np >= 1 && np <= W);

      // ---- blank-only frame: the candidate set is {blank} (a confident blank: most frames of a trained CTC model).
      // Every prefix then keeps exactly its blank-ended entry, scored tot + r_blank; nothing is proposed, the ranking
      // does not change.  If none falls below the relative cut and no two share a history key, the general path
      // would copy every prefix to the same slot of the next generation: update the scores in place instead and
      // leave the hash table alone.  Otherwise (rare) the general path below does the frame.
      const bool blank_only = nc1 == 0 && F.rl[0] != -INFINITY;
      // ---- repeat-only frame: the candidate set is one label c, blank is not in it (the 2nd, 3rd, .. frame of a
      // confidently emitted label).  If every live prefix ends in c and none has blank-ended mass (the previous frame
      // had no blank candidate either), the general path gives every prefix exactly its char-ended entry back, scored
      // snb + r_c: the blank entry is -inf, the extension P+c needs blank-ended mass and is -inf, the parent's
      // contribution is the parent's blank-ended mass and is -inf, and lse2(x, -inf) = x.  Same in-place update.
      const bool repeat_cand = nc1 == 1 && F.rl[0] == -INFINITY;
      // "No two share a history key": every entry of such a frame has the same last char, so two keys are equal iff the
      // prefixes' (closed words, open word) parts are -- a property of the prefix SET, which in-place frames do not
      // change.  The O(np^2) comparison (a fifth of the kernel's instructions on confident input) is therefore made
      // once per generation, by the first in-place frame after a general one (hk_clean).
      if (blank_only || repeat_cand) {
        const int c1 = repeat_cand ? (int)F.cs[0] : 0;
        const double rb = (double)F.rl[c1];
        double m = neg_inf();
        bool general = false;
        SSD_NO_UNROLL
        for (int p = tid; p < np; p += CONS) {
          double v;
          uint32_t lc;
          if (blank_only) {
            v = lse2(A.sb[p], A.snb[p]) + rb;
            lc = 0u;
          } else {
            if (A.last[p] != c1 || A.sb[p] != -INFINITY) general = true;
            v = A.snb[p] + rb;
            lc = (uint32_t)c1;
          }
          newb[p] = v;
          if constexpr (LM) v += lm_of(p);  // the cut below is on the ranking score
          m = v > m ? v : m;
          if (!hk_clean) sel_hk[p] = hist_key(A.hl[p], A.ho[p], (int)lc, p, false, 0ull);  // phase 5's key of that entry
        }
        m = warp_max_d(m);
        if (lane == 0) sc->wmax[warp] = m;
        if (tid == 0) sc->n_b = 0;
        cons_sync<CONS>();
        double mm = sc->wmax[0];
        for (int w = 1; w < CONS / 32; ++w) mm = sc->wmax[w] > mm ? sc->wmax[w] : mm;
        const double cut = mm + (double)s_prune;
        SSD_NO_UNROLL
        for (int p = tid; p < np; p += CONS) {
          double v = newb[p];
          if constexpr (LM) v += lm_of(p);
          if (!(v >= cut) || v == -INFINITY) general = true;
          if (!hk_clean) {
            const uint64_t hk = sel_hk[p];
            if (!hk_dirty) {
              // through the (zeroed) history-key table: whoever finds its key already there has a twin.  O(np) instead
              // of O(np^2) comparisons -- at width 200 the pairwise scan was a third of a confident utterance's time.
              int s_ = (int)(hk & (uint64_t)(HKS - 1));
              int mine = -1;
              int tries = 0;
              for (; tries < HKS; ++tries) {  // bounded: the table has at least 2 np slots, so this never runs out
                const unsigned long long old = atomicCAS(&hk_keys[s_], 0ull, (unsigned long long)hk);
                if (old == 0ull) { mine = s_; break; }
                if (old == (unsigned long long)hk) { general = true; break; }
                s_ = (s_ + 1) & (HKS - 1);
              }
              if (tries == HKS) general = true;
              sel_slot[p] = mine;
            } else {  // the bucket cut left its scratch in the table's memory: compare pairwise
              SSD_NO_UNROLL_INNER
              for (int q = 0; q < np; ++q)
                if (q != p && sel_hk[q] == hk) general = true;
            }
          }
        }
        if (general) sc->n_b = 1;
        cons_sync<CONS>();
        if (!hk_clean && !hk_dirty) {  // the table goes back to zero: every thread clears the slots it claimed
          SSD_NO_UNROLL
          for (int p = tid; p < np; p += CONS)
            if (sel_slot[p] >= 0) hk_keys[sel_slot[p]] = 0ull;
        }
        if (sc->n_b == 0) {
          hk_clean = true;  // in particular no two keys were equal; holds until a general frame rebuilds the prefixes
          SSD_NO_UNROLL
          for (int p = tid; p < np; p += CONS) {
            if (blank_only) {
              A.sb[p] = newb[p];
              A.snb[p] = neg_inf();
            } else {
              A.snb[p] = newb[p];
            }
          }
#ifdef SSD_BEAM_PROF
          const long long pf_t1 = clock64();
          pf_fast += pf_t1 - pf_t0;
          ++pf_nfast;
#endif
          bar_sync(BAR_FRAME, THREADS);  // same generation, same table; frame t+1's record is visible
#ifdef SSD_BEAM_PROF
          pf_wait += clock64() - pf_t1;
#endif
          continue;
        }
      }

#ifdef SSD_BEAM_PROF
      pf_last = clock64();
#endif
      // ---- phase 1: tot, parent slot.  (When no staircase row is cut -- see phase 2 -- nobody needs another prefix's
      // tot before phase 3 except for the parent's, which is then recomputed on the spot: phases 1 and 2 share a barrier.)
      const bool cut_rows = !LM && nc1 > W / np + 1;  // under LM scores the staircase bound does not hold: full rows
      SSD_NO_UNROLL
      for (int p = tid; p < np; p += CONS) {
        tot[p] = lse2(A.sb[p], A.snb[p]);
        par[p] = table_find(tkeys, tvals, tmask, A.Hpar[p]);
      }
      if (cut_rows) cons_sync<CONS>();

      PF_STAMP(0);
      // ---- phase 2: rank by tot; per-prefix entries.  The rank only sets the staircase width of a row; when even the
      // last row keeps all nc1 candidates (few candidates: every frame of a confident model) no row is cut and the
      // prefixes keep their order.
      const float r_blank = F.rl[0];
      SSD_NO_UNROLL
      for (int p = tid; p < np; p += CONS) {
        const double mine = tot[p];
        int rk = p;
        if (cut_rows) {
          rk = 0;
          SSD_NO_UNROLL_FLAT
          for (int q = 0; q < np; ++q) {
            const double o = tot[q];
            rk += (o > mine || (o == mine && q < p)) ? 1 : 0;
          }
        }
        rank2p[rk] = p;
        const int e = A.last[p];
        const float r_e = (e != LBL_NONE) ? F.rl[e] : -INFINITY;
        newb[p] = (r_blank != -INFINITY) ? mine + (double)r_blank : neg_inf();
        double rep = neg_inf(), ext = neg_inf();
        if (r_e != -INFINITY) {
          // repeat of the last char keeps the text; an absorbed space also collects the blank-ended mass
          rep = (e == space_lbl ? mine : A.snb[p]) + (double)r_e;
          const int pp = par[p];
          if (pp >= 0) ext = ((A.last[pp] != e) ? (cut_rows ? tot[pp] : lse2(A.sb[pp], A.snb[pp])) : A.sb[pp]) + (double)r_e;
        }
        newnb[p] = lse2(rep, ext);
        // staircase width of this row (no row cut: every row holds nc1 proposals, its offset is known without a scan)
        int J = W / (rk + 1) + 1;
        rowoff[rk] = cut_rows ? (J < nc1 ? J : nc1) : rk * nc1;
      }
      cons_sync<CONS>();

      PF_STAMP(1);
      // ---- phase 2b: exclusive scan of row widths (warp 0), only if some row is cut
      if (cut_rows) {
      if (warp == 0) {
        int carry = 0;
        SSD_NO_UNROLL_INNER
        for (int i0 = 0; i0 < np; i0 += 32) {
          const int i = i0 + lane;
          const int v = i < np ? rowoff[i] : 0;
          int incl = v;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += y;
          }
          if (i < np) rowoff[i] = carry + incl - v;
          carry += __shfl_sync(FULL, incl, 31);
        }
        if (lane == 0) sc->n_e3 = carry;
      }
      cons_sync<CONS>();
      }

      PF_STAMP(2);
      // ---- phase 3: candidate entries (and, on the way, their maximum: phase 4a)
      const int n_e3 = cut_rows ? sc->n_e3 : np * nc1;
      const int ncand = 2 * np + n_e3;
      SSD_ASSERT(ncand <= NCAND);
      // descriptor of candidate k: (prefix << 8) | label, label 0 = blank-ended entry, 0xFE = char-ended entry.
      // The LM kernel's rows are uniform, so it derives the descriptor from the index instead of storing it.
      auto desc_of = [&](int k) -> uint32_t {
        if constexpr (LM) {
          if (k < np) return (uint32_t)k << 8;
          if (k < 2 * np) return ((uint32_t)(k - np) << 8) | 0xFEu;
          const int q = k - 2 * np, row = q / nc1;
          return ((uint32_t)row << 8) | (uint32_t)F.cs[q - row * nc1];
        } else {
          return cdesc[k];
        }
      };
      double m_loc = neg_inf();
      constexpr int P3U = LM ? SSD_BEAM_LM_P3U : 1;
      (void)P3U;
#ifdef SSD_EMU
      for (int k = tid; k < ncand; k += CONS) {
#else
#pragma unroll(P3U)  // LM: a proposal waits for a trie / n-gram load from L2; unrolling keeps several in flight
      for (int k = tid; k < ncand; k += CONS) {
#endif
        double s;
        uint32_t d;
        if (k < np) {
          s = newb[k];
          d = ((uint32_t)k << 8) | 0u;  // blank-ended entry of prefix k
        } else if (k < 2 * np) {
          const int p = k - np;
          s = newnb[p];
          d = ((uint32_t)p << 8) | 0xFEu;  // char-ended entry of prefix p
        } else {
          const int q = k - 2 * np;
          int lo, col;
          if ((LM || SSD_BEAM_ROWDIV) && !cut_rows) {  // uniform rows of nc1 proposals in prefix order
            lo = q / nc1;
            col = q - lo * nc1;
          } else {
            lo = 0;
            int hi = np - 1;  // last row with rowoff <= q
            while (lo < hi) {
              const int mid = (lo + hi + 1) >> 1;
              if (rowoff[mid] <= q) lo = mid; else hi = mid - 1;
            }
            col = q - rowoff[lo];
          }
          const int p = cut_rows ? rank2p[lo] : lo;
          const int c = (int)F.cs[col];
          const int e = A.last[p];
          s = neg_inf();
          if (!(c == space_lbl && e == space_lbl)) {  // absorbed space is not an extension
            const uint64_t h = hmix(A.H[p], (uint32_t)c);
            if (table_find(tkeys, tvals, tmask, h) < 0)  // else folded into the live child's snb
              s = ((c == e) ? A.sb[p] : tot[p]) + (double)F.rl[c];
            if constexpr (LM) {
              // upstream LM-scores every merged beam before pruning: the closed text enters its score cache
              if (c == space_lbl) seen_insert(seen, seen_mask, h);
            }
          }
          d = ((uint32_t)p << 8) | (uint32_t)c;
        }
        if constexpr (!LM) {
          cscore[k] = s;
          cdesc[k] = d;
        }
        if constexpr (LM) {
          if (s != -INFINITY) {
            const int p = (int)(d >> 8), c = (int)(d & 0xFFu);
            double lm;
            if (c == 0 || c == 0xFE) {
              lm = lm_of(p);
            } else if (c == space_lbl) {  // closes the open word: score it as a word after the prefix's LM state
              const int node = trieA[p];
              lm = lmtA[p] + lm_word_score(P.lm, P.lm_factor, ctxA + p * CX, node >= 0 ? P.lm.trie_word[node] : -1, false, nullptr);
            } else {
              const int node = trieA[p];
              const int child = node >= 0 ? P.lm.trie_child[(int64_t)node * P.lm.n_labels + c] : -1;
              lm = lmtA[p] + lm_partial_score(P.lm, child, plenA[p] + 1);
            }
            s = s + lm;
          }
          cfused[k] = s;
        }
        m_loc = s > m_loc ? s : m_loc;
      }
      m_loc = warp_max_d(m_loc);
      if (lane == 0) sc->wmax[warp] = m_loc;
      cons_sync<CONS>();

      PF_STAMP(3);
      // ---- phase 4a: block max (the warps' maxima were left in sc->wmax by phase 3)
      double cutoff;
      {
        double m = sc->wmax[0];
        for (int w = 1; w < CONS / 32; ++w) m = sc->wmax[w] > m ? sc->wmax[w] : m;
        cutoff = m + (double)s_prune;
      }
      PF_STAMP(4);
      // ---- phase 4b: compact entries that survive the relative prune (order preserved)
      int n_alive;
      {
        // ballots over chunks of CONS.  (One scan over per-thread runs of candidates -- two barriers however many
        // candidates -- measured 4 % slower on config 3: the scan's five dependent shuffles cost more than the barriers.)
        int running = 0, buf = 0;
        SSD_NO_UNROLL
        for (int base_k = 0; base_k < ncand; base_k += CONS) {
          const int k = base_k + tid;
          const bool alive = k < ncand && rscore[k] >= cutoff && rscore[k] != -INFINITY;
          const unsigned m = __ballot_sync(FULL, alive);
          if (lane == 0) sc->wcnt[buf][warp] = __popc(m);
          cons_sync<CONS>();
          int off = running, totc = 0;
#pragma unroll
          for (int w = 0; w < CONS / 32; ++w) {
            const int c = sc->wcnt[buf][w];
            if (w < warp) off += c;
            totc += c;
          }
          if (alive) {
            const int pos = off + __popc(m & ((1u << lane) - 1u));
            aidx[pos] = (ai_t)k;
            if (pos < W) sel[pos] = k;
          }
          running += totc;
          buf ^= 1;  // the next chunk's counts go to the other buffer: its barrier separates this read from the next write here
        }
        n_alive = running;
      }
      cons_sync<CONS>();
      const int n_sel = n_alive < W ? n_alive : W;
      PF_STAMP(5);
      // ---- phase 4c: the W best alive entries go to sel[rank].
      // More than W alive (flat posteriors): first cut down to exactly W with a 256-bucket histogram
      // over [cutoff, max] -- everything in buckets above the boundary bucket is in, the boundary
      // bucket is resolved by exact rank among its few members -- then rank only those W.
      // Same ordering rule as a full rank count (score desc, candidate index asc), so same result.
      // (At most W alive: all of them are selected, in candidate order -- the order of the selection only decides
      // slot and node numbers, and the best entry per history key is found by score below, not by rank.)
      if (n_alive > W) {
        hk_dirty = true;  // the bucket cut works in the memory the history-key table lives in
        SSD_NO_UNROLL
        for (int i = tid; i < 256; i += CONS) hist[i] = 0;
        if (tid == 0) sc->n_b = 0;
        cons_sync<CONS>();
        const double top = cutoff - (double)s_prune;
        const double scale = s_prune < 0.f ? (double)(BK_TOP + 1) / -(double)s_prune : 0.0;
        SSD_NO_UNROLL
        for (int i = tid; i < n_alive; i += CONS) {
          int b = (int)((top - rscore[aidx[i]]) * scale);
          b = b < 0 ? 0 : (b > BK_TOP ? BK_TOP : b);
          abkt[i] = (bk_t)(b + BK_OFF);
          atomicAdd(&hist[b], 1);
        }
        cons_sync<CONS>();
        if (warp == 0) {  // boundary bucket: first one where the running count reaches W
          int c[8], tot = 0;
#pragma unroll
          for (int k = 0; k < 8; ++k) { c[k] = hist[lane * 8 + k]; tot += c[k]; }
          int incl = tot;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += y;
          }
          int run = incl - tot;  // entries in buckets before my 8
          if (run < W && incl >= W) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (run < W && run + c[k] >= W) { sc->beta = lane * 8 + k; sc->c0 = run; }
              run += c[k];
            }
          }
        }
        cons_sync<CONS>();
        const int beta = sc->beta, need = W - sc->c0;
        SSD_NO_UNROLL
        for (int i = tid; i < n_alive; i += CONS)
          if ((int)abkt[i] == beta + BK_OFF) blist[atomicAdd(&sc->n_b, 1)] = (ai_t)i;
        cons_sync<CONS>();
        const int m = sc->n_b;
        SSD_NO_UNROLL
        for (int t = tid; t < m; t += CONS) {
          const int i = blist[t];
          const double mine = rscore[aidx[i]];
          int rk = 0;
          SSD_NO_UNROLL_FLAT
          for (int u = 0; u < m; ++u) {
            const int j = blist[u];
            const double o = rscore[aidx[j]];
            rk += (o > mine || (o == mine && j < i)) ? 1 : 0;
          }
          abkt[i] = (bk_t)(rk < need ? BK_IN : BK_OUT);  // in / out
        }
        cons_sync<CONS>();
        // compact the W survivors, alive order preserved
        {
          int running = 0;
          SSD_NO_UNROLL
          for (int base_i = 0; base_i < n_alive; base_i += CONS) {
            const int i = base_i + tid;
            const bool in = i < n_alive && (int)abkt[i] < beta + BK_OFF;
            const unsigned mk = __ballot_sync(FULL, in);
            if (lane == 0) sc->wcnt[0][warp] = __popc(mk);
            cons_sync<CONS>();
            int off = running, totc = 0;
#pragma unroll
            for (int w = 0; w < CONS / 32; ++w) {
              const int cw = sc->wcnt[0][w];
              if (w < warp) off += cw;
              totc += cw;
            }
            if (in) {
              const int pos = off + __popc(mk & ((1u << lane) - 1u));
              bidx[pos] = aidx[i];
            }
            running += totc;
            cons_sync<CONS>();
          }
          SSD_ASSERT(running == W);
        }
        SSD_NO_UNROLL
        for (int i = tid; i < W; i += CONS) sel[i] = bidx[i];
        cons_sync<CONS>();  // (at most W alive: sel[] was complete behind phase 4b's barrier)
      }

      PF_STAMP(6);
      // ---- phase 5: history keys, then keep the best entry per key.
      // Best entry per key through a small hash table (over scratch the cut above no longer needs): highest score, and
      // among equal scores the lowest candidate index -- what "first in rank order" means -- found with one atomicMax
      // and one atomicMin per entry.  Replaces a full ranking of the selection followed by a scan of every
      // better-ranked entry (a quarter of the kernel's instructions at config 3).  The table is kept ZEROED between
      // uses (re-zeroed in phase 6a, and here only after the cut has used its memory), so the keys are inserted by
      // the same loop that computes them: three barriers for the phase instead of five.
      const bool small_sel = n_sel <= P.small_sel;  // a handful of entries (narrow beams): compare directly, one pass
      if (hk_dirty && !small_sel) {
        SSD_NO_UNROLL
        for (int i = tid; i < HKS; i += CONS) {
          hk_keys[i] = 0ull;
          hk_best[i] = 0ull;
          hk_idx[i] = 0x7fffffff;
        }
        cons_sync<CONS>();
      }
      SSD_NO_UNROLL
      for (int r = tid; r < n_sel; r += CONS) {
        const uint32_t d = desc_of(sel[r]);
        const int p = (int)(d >> 8), c = (int)(d & 0xFFu);
        uint64_t hl = A.hl[p], ho = A.ho[p];
        int lc;
        bool shifted = false;
        if (c == 0) lc = 0;
        else if (c == 0xFE) lc = A.last[p];
        else {
          lc = c;
          if (c == space_lbl) { hl = ho; ho = 0ull; shifted = true; }
          else ho = hmix(ho, (uint32_t)c);
        }
        // one 64-bit key for (last word(s), open word, last char)
        const uint64_t hk = hist_key(hl, ho, lc, p, shifted, A.hl[p]);
        sel_hk[r] = hk;
        if (!small_sel) {
          int s_ = (int)(hk & (uint64_t)(HKS - 1));
          for (;;) {
            const unsigned long long old = atomicCAS(&hk_keys[s_], 0ull, (unsigned long long)hk);
            if (old == 0ull || old == (unsigned long long)hk) break;
            s_ = (s_ + 1) & (HKS - 1);
          }
          atomicMax(&hk_best[s_], order_key(rscore[sel[r]]));
          sel_slot[r] = s_;  // remembered for the two passes below (sel_slot is rewritten in phase 6a)
        }
      }
      SSD_NO_UNROLL
      for (int p = tid; p < np; p += CONS) keepf[p] = 0;
      cons_sync<CONS>();
      if (!small_sel) {
        SSD_NO_UNROLL
        for (int r = tid; r < n_sel; r += CONS) {
          const int s_ = sel_slot[r];
          if (order_key(rscore[sel[r]]) == hk_best[s_]) atomicMin(&hk_idx[s_], sel[r]);
        }
        cons_sync<CONS>();
      }
      SSD_NO_UNROLL
      for (int r = tid; r < n_sel; r += CONS) {
        int keep;
        if (small_sel) {
          const uint64_t hk = sel_hk[r];
          const double mine = rscore[sel[r]];
          keep = 1;
          SSD_NO_UNROLL_INNER
          for (int q = 0; q < n_sel; ++q) {
            if (q == r || sel_hk[q] != hk) continue;
            const double o = rscore[sel[q]];
            if (o > mine || (o == mine && sel[q] < sel[r])) keep = 0;
          }
        } else {
          keep = hk_idx[sel_slot[r]] == sel[r] ? 1 : 0;
        }
        const uint32_t d = desc_of(sel[r]);
        const int c = (int)(d & 0xFFu);
        if (keep && (c == 0 || c == 0xFE)) {
          atomicOr(&keepf[d >> 8], c == 0 ? 1 : 2);
          keep = 0;  // lives on inside an existing prefix, no new slot
          sel_keep[r] = 0;
        } else {
          sel_keep[r] = keep;
        }
      }
      cons_sync<CONS>();

      PF_STAMP(7);
      // ---- phase 6a: slots.  warp 0 scans surviving old prefixes, warp 1 the new ones.
      if (warp == 0) {
        int carry = 0;
        SSD_NO_UNROLL_INNER
        for (int i0 = 0; i0 < np; i0 += 32) {
          const int i = i0 + lane;
          const bool f = i < np && keepf[i] != 0;
          const unsigned m = __ballot_sync(FULL, f);
          if (i < np) oslot[i] = f ? carry + __popc(m & ((1u << lane) - 1u)) : -1;
          carry += __popc(m);
        }
        if (lane == 0) sc->n_keep_old = carry;
      }
      if (warp == (CONS > 32 ? 1 : 0)) {
        int carry = 0;
        SSD_NO_UNROLL_INNER
        for (int i0 = 0; i0 < n_sel; i0 += 32) {
          const int i = i0 + lane;
          const bool f = i < n_sel && sel_keep[i] != 0;
          const unsigned m = __ballot_sync(FULL, f);
          if (i < n_sel) sel_slot[i] = f ? carry + __popc(m & ((1u << lane) - 1u)) : -1;
          carry += __popc(m);
        }
        if (lane == 0) sc->n_new = carry;
      }
      SSD_NO_UNROLL
      for (int i = tid; i < TS; i += CONS) tkeys[i] = 0ull;
      if (!small_sel) {  // the history-key table goes back to zero for its next use (its last read is behind the barrier above)
        SSD_NO_UNROLL
        for (int i = tid; i < HKS; i += CONS) {
          hk_keys[i] = 0ull;
          hk_best[i] = 0ull;
          hk_idx[i] = 0x7fffffff;
        }
        hk_dirty = false;  // (a small selection leaves the table as it is: dirty stays dirty)
      }
      cons_sync<CONS>();

      PF_STAMP(8);
      // ---- phase 6b: write the next generation and its hash table
      const int n_old = sc->n_keep_old, n_new = sc->n_new;
      SSD_ASSERT(n_old + n_new >= 1 && n_old + n_new <= W);
      SSD_NO_UNROLL
      for (int p = tid; p < np; p += CONS) {
        const int s = oslot[p];
        if (s < 0) continue;
        N.H[s] = A.H[p]; N.Hpar[s] = A.Hpar[p]; N.Hstrip[s] = A.Hstrip[p];
        N.hl[s] = A.hl[p]; N.ho[s] = A.ho[p];
        N.node[s] = A.node[p]; N.last[s] = A.last[p]; N.len[s] = A.len[p];
        N.sb[s] = (keepf[p] & 1) ? newb[p] : neg_inf();
        N.snb[s] = (keepf[p] & 2) ? newnb[p] : neg_inf();
        table_insert(tkeys, tvals, tmask, A.H[p], s);
        if constexpr (LM) {
          lmtN[s] = lmtA[p]; trieN[s] = trieA[p]; plenN[s] = plenA[p];
          for (int j = 0; j < CX; ++j) { ctxN[s * CX + j] = ctxA[p * CX + j]; hwN[s * CX + j] = hwA[p * CX + j]; }
        }
      }
      SSD_NO_UNROLL
      for (int r = tid; r < n_sel; r += CONS) {
        const int j = sel_slot[r];
        if (j < 0) continue;
        const int s = n_old + j;
        const int k = sel[r];
        const uint32_t d = desc_of(k);
        const int p = (int)(d >> 8), c = (int)(d & 0xFFu);
        const uint64_t h = hmix(A.H[p], (uint32_t)c);
        N.H[s] = h; N.Hpar[s] = A.H[p];
        if (c == space_lbl) { N.Hstrip[s] = A.H[p]; N.hl[s] = A.ho[p]; N.ho[s] = 0ull; }
        else { N.Hstrip[s] = h; N.hl[s] = A.hl[p]; N.ho[s] = hmix(A.ho[p], (uint32_t)c); }
        N.last[s] = c; N.len[s] = A.len[p] + 1;
        // the beam carries the ACOUSTIC score on; the LM kernel kept only the ranking score and recomputes it (a
        // selected proposal is never a folded or absorbed one: those rank -inf)
        N.sb[s] = neg_inf();
        N.snb[s] = LM ? ((c == A.last[p]) ? A.sb[p] : tot[p]) + (double)F.rl[c] : cscore[k];
        const unsigned nd = node0 + (unsigned)j;
        SSD_ASSERT((int64_t)nd < P.node_cap);
        N.node[s] = nd;
        nodes[nd] = make_uint2(A.node[p], (unsigned)c);
        table_insert(tkeys, tvals, tmask, h, s);
        if constexpr (LM) {
          if (c == space_lbl) {  // the open word is closed: its score joins the text's, the LM state moves on
            const int node = trieA[p];
            int st[LM_CX];
            lmtN[s] = lmtA[p] + lm_word_score(P.lm, P.lm_factor, ctxA + p * CX, node >= 0 ? P.lm.trie_word[node] : -1, false, st);
            trieN[s] = 0; plenN[s] = 0;
            for (int j = 0; j < CX; ++j) {
              ctxN[s * CX + j] = j < P.lm.order - 1 ? st[j] : -1;
              hwN[s * CX + j] = j == 0 ? A.hl[p] : hwA[p * CX + j - 1];
            }
          } else {
            const int node = trieA[p];
            lmtN[s] = lmtA[p];
            trieN[s] = node >= 0 ? P.lm.trie_child[(int64_t)node * P.lm.n_labels + c] : -1;
            plenN[s] = plenA[p] + 1;
            for (int j = 0; j < CX; ++j) { ctxN[s * CX + j] = ctxA[p * CX + j]; hwN[s * CX + j] = hwA[p * CX + j]; }
          }
        }
      }
      if (tid == 0) {
        sc->np = n_old + n_new;
        sc->nodes_used = node0 + (unsigned)n_new;
      }
      g ^= 1;
      hk_clean = false;
      PF_STAMP(9);
#ifdef SSD_BEAM_PROF
      const long long pf_t1 = clock64();
      pf_gen += pf_t1 - pf_t0;
      ++pf_ngen;
#endif
      bar_sync(BAR_FRAME, THREADS);  // next generation, table, and frame t+1 record are all visible
#ifdef SSD_BEAM_PROF
      pf_wait += clock64() - pf_t1;
#endif
    }
#ifdef SSD_BEAM_PROF
    if (!producer && blockIdx.x == 0 && tid == 0)
    {
      printf("consumers: %d fast frames at %lld cycles, %d general frames at %lld cycles, waiting %lld cycles/frame\n", pf_nfast,
             pf_nfast ? pf_fast / pf_nfast : 0, pf_ngen, pf_ngen ? pf_gen / pf_ngen : 0, len ? pf_wait / len : 0);
      const int n = pf_ngen ? pf_ngen : 1;
      printf("  general frame by phase (cycles): 1 %lld | 2 %lld | 2b %lld | 3 %lld | 4a %lld | 4b %lld | 4c %lld | 5 %lld | 6a %lld | 6b %lld\n",
             pf_ph[0] / n, pf_ph[1] / n, pf_ph[2] / n, pf_ph[3] / n, pf_ph[4] / n, pf_ph[5] / n, pf_ph[6] / n, pf_ph[7] / n, pf_ph[8] / n, pf_ph[9] / n);
    }
#endif

    // ---------------- finish: merge variants, pick the best text, emit labels
    __syncthreads();
    if (!producer) {
      const PrefixView A = prefix_view(gen0, g, gs8, gs4);
      const int np = sc->np;
      SSD_NO_UNROLL
      for (int p = tid; p < np; p += CONS) tot[p] = lse2(A.sb[p], A.snb[p]);
      cons_sync<CONS>();
      SSD_NO_UNROLL
      for (int p = tid; p < np; p += CONS) {
        // group = prefixes with equal text after dropping a trailing space
        const uint64_t hs = A.Hstrip[p];
        double f = neg_inf();
        int rep = p;
        SSD_NO_UNROLL_INNER
        for (int q = 0; q < np; ++q)
          if (A.Hstrip[q] == hs) {
            f = lse2(f, tot[q]);
            if (q < rep) rep = q;
          }
        if constexpr (LM) {
          // upstream's last pass: the open word is closed and scored WITH the end-of-sentence term -- unless the
          // resulting text is already in its score cache (it was scored as a closed text in some frame), whose
          // entry has none; texts that end closed are always in the cache.
          if (rep == p) {
            const double* lmtF = lmt_base + g * lm_gs8;
            const int* trieF = ltrie_base + g * lm_gs4;
            const int* plenF = lplen_base + g * lm_gs4;
            const int* ctxF = lctx_base + g * lm_gs4;
            int q_open = -1, q_closed = -1;
            SSD_NO_UNROLL_INNER
            for (int q = 0; q < np; ++q)
              if (A.Hstrip[q] == hs) {
                if (plenF[q] > 0) q_open = q; else q_closed = q;
              }
            double lm;
            if (P.lm.eos_rescore) {
              // the other published variant (score cache keyed by (text, is_eos)): every final beam is re-scored with
              // the </s> term.  Upstream's final merge leaves the (text, last word) split of the LAST of the merged
              // beams, i.e. of the group's lowest-ranked entry: the open prefix scores its open word; the closed one an
              // empty word, which kenlm takes for <unk> and the decoder for out of vocabulary.
              double eo = INFINITY, ec = INFINITY;
              if (q_open >= 0) {
                const double l = lmtF[q_open] + lm_partial_score(P.lm, trieF[q_open], plenF[q_open]);
                if (A.sb[q_open] != -INFINITY) eo = A.sb[q_open] + l;
                if (A.snb[q_open] != -INFINITY && A.snb[q_open] + l < eo) eo = A.snb[q_open] + l;
              }
              if (q_closed >= 0) {
                const double l = lmtF[q_closed];
                if (A.sb[q_closed] != -INFINITY) ec = A.sb[q_closed] + l;
                if (A.snb[q_closed] != -INFINITY && A.snb[q_closed] + l < ec) ec = A.snb[q_closed] + l;
              }
              if (q_open >= 0 && (q_closed < 0 || eo <= ec)) {
                const int node = trieF[q_open];
                lm = lmtF[q_open] + lm_word_score(P.lm, P.lm_factor, ctxF + q_open * CX, node >= 0 ? P.lm.trie_word[node] : -1, true, nullptr);
              } else {
                lm = lmtF[q_closed] + lm_word_score(P.lm, P.lm_factor, ctxF + q_closed * CX, -1, true, nullptr);
              }
            } else if (q_open >= 0) {
              const bool cached = space_lbl >= 0 && seen_find(seen, seen_mask, hmix(A.H[q_open], (uint32_t)space_lbl));
              const int node = trieF[q_open];
              lm = lmtF[q_open] + lm_word_score(P.lm, P.lm_factor, ctxF + q_open * CX, node >= 0 ? P.lm.trie_word[node] : -1, !cached, nullptr);
            } else {
              lm = lmtF[q_closed];
            }
            f = f + lm;
          }
        }
        newb[p] = (rep == p) ? f : neg_inf();  // one representative per group competes
      }
      cons_sync<CONS>();
      if (tid == 0) {
        int best = 0;
        SSD_NO_UNROLL_INNER
        for (int p = 1; p < np; ++p)
          if (newb[p] > newb[best]) best = p;
        int n = A.len[best];
        unsigned nd = A.node[best];
        int32_t* out = P.out_tok + ob * P.T;
        if (n > 0 && A.last[best] == space_lbl) {  // trailing space is not part of the text
          nd = nodes[nd].x;
          --n;
        }
        P.out_len[ob] = n;
        if (P.out_score) P.out_score[ob] = len > 0 ? newb[best] : 0.0;
        for (int i = n - 1; i >= 0; --i) {
          const uint2 rec = nodes[nd];
          out[i] = lab2voc[rec.y];
          nd = rec.x;
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ host side
// The instantiations: capacity as a constant for the reference's widths (config_builder.py:68-93: 20 / 50 / 100), run-time
// otherwise.  SSD_BEAM_PICK(lm, width, F) evaluates F(kernel) with the instantiation serving that launch.
#ifndef SSD_BEAM_WCAP
#define SSD_BEAM_WCAP 1
#endif
#if SSD_BEAM_WCAP
#define SSD_BEAM_PICK(lm, width, F)                                                 \
  ((lm) ? F((beam_kernel<true, 0>)) /* measured: no gain for the LM kernel (67.5 k against 66.6 k), one instance */ \
        : ((width) == 50 ? (vl47 ? F((beam_kernel<false, 50, 47>)) : F((beam_kernel<false, 50>)))  \
           : (width) == 20 ? (vl47 ? F((beam_kernel<false, 20, 47>)) : F((beam_kernel<false, 20>)))    \
           : (width) == 100 ? (vl47 ? F((beam_kernel<false, 100, 47>)) : F((beam_kernel<false, 100>))) \
           : (width) == 1 ? F((beam_kernel<false, 1>))                                \
           : (width) == 10 ? F((beam_kernel<false, 10>))                              \
           : (width) == 200 ? F((beam_kernel<false, 200>)) : F((beam_kernel<false, 0>))))
#else
#define SSD_BEAM_PICK(lm, width, F) ((void)vl47, (lm) ? F((beam_kernel<true, 0>)) : F((beam_kernel<false, 0>)))
#endif

static int beam_grid(int64_t B, size_t smem, bool lm, int width, int vl) {
  const bool vl47 = vl == 47;
  (void)vl47;
#ifdef SSD_EMU
  (void)smem; (void)lm; (void)width;
  return (int)(B < 3 ? B : 3);
#else
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define SSD_F_OCC(K) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, K, lm ? BM_LM_THREADS : BM_THREADS, smem)
  SSD_BEAM_PICK(lm, width, SSD_F_OCC);
#undef SSD_F_OCC
  if (per_sm < 1) per_sm = 1;
  const int64_t g = (int64_t)sms * per_sm;
  return (int)(B < g ? B : g);
#endif
}

static int beam_vl(int64_t V, int blank, int pad) { return (int)V - 1 - ((pad >= 0 && pad < V && pad != blank) ? 1 : 0) + 1; }

static int64_t beam_node_cap(int64_t T, int W) { return T * (int64_t)W + 2; }

static int beam_max_grid() {
#ifdef SSD_EMU
  return 3;
#else
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms * 12;
#endif
}

// Resident CTAs of a launch over `items` work items at capacity `width`: what the occupancy calculator gives for the
// launch's shared-memory footprint, never more than there are items.  The workspace is sized from the same number
// (the footprint is taken with the smaller of the two possible label counts, so it can only over-estimate the grid).
static int beam_grid_for(int64_t items, int64_t V, int VL, int width, size_t* smem_out, int lm_order = 0) {
  const int TS = beam_table_slots(width), NCAND = lm_order > 0 ? beam_ncand_cap_lm(width, VL) : beam_ncand_cap(width, VL);
  const size_t smem = beam_smem_bytes(width, (int)V, VL, TS, NCAND, lm_order);
  if (smem_out) *smem_out = smem;
  if (smem > 227 * 1024) return 0;
  int grid = beam_grid(items, smem, lm_order > 0, width, VL);
  const int gmax = beam_max_grid();
  return grid > gmax ? gmax : grid;
}

size_t beam_workspace_bytes(int64_t B, int64_t T, int64_t V, int width) {
  int g = beam_grid_for(B, V, (int)V - 1, width, nullptr);
  if (g < 1) g = 1;
  return 256 + (size_t)g * (size_t)beam_node_cap(T, width) * sizeof(uint2);
}

// slots of one CTA's set of scored closed texts: a power of two >= 2 * T * W
static int64_t beam_seen_cap(int64_t T, int W) {
  int64_t n = 1024;
  while (n < 2 * T * (int64_t)W) n <<= 1;
  return n;
}
static size_t beam_nodes_bytes(int grid, int64_t T, int W) {
  return ((size_t)grid * (size_t)beam_node_cap(T, W) * sizeof(uint2) + 255) & ~(size_t)255;
}
// LM launches: [counter 256 B | back-pointer nodes | scored-text sets]; the grid is bounded with the smallest
// per-prefix LM state (order 1), which can only over-estimate it
size_t beam_lm_workspace_bytes(int64_t B, int64_t T, int64_t V, int width) {
  int g = beam_grid_for(B, V, (int)V - 1, width, nullptr, 1);
  if (g < 1) g = 1;
  return 256 + beam_nodes_bytes(g, T, width) + (size_t)g * (size_t)beam_seen_cap(T, width) * sizeof(unsigned long long);
}

int beam_launch(const float* lp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st, const int64_t* lengths,
                const int32_t* order, int blank, int pad, int space, const ssd_beam_setting* settings, int n_set,
                const ssd_lm_tables* lm, int32_t* out_tok, int32_t* out_len, double* out_score, void* ws, size_t ws_bytes,
                cudaStream_t stream) {
  if (B == 0 || n_set == 0) return SSD_OK;
  BeamParams P;
  P.lm = ssd_lm_tables();
  P.lm_factor = 1.0 / std::log10(std::exp(1.0));
  P.seen = nullptr;
  P.seen_cap = 0;
  if (lm) P.lm = *lm;
  const int lm_order = lm ? lm->order : 0;
  P.lp = lp; P.T = T; P.sb = sb; P.st = st; P.B = (int)B; P.V = (int)V;
  P.lengths = lengths; P.order = order;
  P.blank = blank; P.pad = (pad >= 0 && pad < V) ? pad : -1; P.space = space;
  int width = 0;
  P.n_set = n_set;
  for (int i = 0; i < n_set; ++i) {
    P.set_w[i] = settings[i].beam_width;
    P.set_bias[i] = settings[i].blank_bias;
    P.set_prune[i] = settings[i].beam_prune_logp;
    P.set_tmin[i] = settings[i].token_min_logp;
    width = settings[i].beam_width > width ? settings[i].beam_width : width;
  }
  P.W = width;
  {
    const char* e = std::getenv("SSD_BEAM_SMALL");
    P.small_sel = e && *e ? std::atoi(e) : 8;  // measured: 24 costs 4 % on config 3, 0 costs 10 % at width 1
  }
  P.out_tok = out_tok; P.out_len = out_len; P.out_score = out_score;
  P.VL = beam_vl(V, blank, P.pad);
  P.TS = beam_table_slots(width);
  P.NCAND = lm ? beam_ncand_cap_lm(width, P.VL) : beam_ncand_cap(width, P.VL);
  if (lm && lm->n_labels != P.VL)
    return fail(SSD_ERR_BAD_ARG, "beam: the LM's character trie is laid out for %d labels, the decoder has %d", lm->n_labels, P.VL);
  P.use_tma = (st == V) && aligned16(lp) && ((sb * 4) % 16 == 0) && ((V * 4) % 16 == 0) ? 1 : 0;
  size_t smem = 0;
  const int grid = beam_grid_for(B * n_set, V, P.VL, width, &smem, lm_order);
  if (grid < 1) return fail(SSD_ERR_UNSUPPORTED, "beam: beam_width %d needs %zu bytes of shared memory (> 227 KB)", width, smem);
  const bool vl47 = P.VL == 47;
  (void)vl47;
#ifndef SSD_EMU
#define SSD_F_ATTR(K) cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
  cudaError_t e = SSD_BEAM_PICK(lm != nullptr, width, SSD_F_ATTR);
#undef SSD_F_ATTR
  if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "beam: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
#endif
  P.node_cap = beam_node_cap(T, width);
  size_t need = 256 + (size_t)grid * (size_t)P.node_cap * sizeof(uint2);
  if (lm) {
    P.seen_cap = beam_seen_cap(T, width);
    need = 256 + beam_nodes_bytes(grid, T, width) + (size_t)grid * (size_t)P.seen_cap * sizeof(unsigned long long);
  }
  if (!ws || ws_bytes < need) return fail(SSD_ERR_WORKSPACE, "beam: workspace %zu bytes, need %zu", ws_bytes, need);
  P.counter = reinterpret_cast<unsigned*>(ws);
  P.nodes = reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(ws) + 256);
  if (lm) P.seen = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(ws) + 256 + beam_nodes_bytes(grid, T, width));
#ifdef SSD_EMU
  *P.counter = 0u;
#else
  e = cudaMemsetAsync(P.counter, 0, 256, stream);
  if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "beam: cudaMemsetAsync: %s", cudaGetErrorString(e));
#endif
#define SSD_F_GO(K) (SSD_LAUNCH(K, (unsigned)grid, (lm ? BM_LM_THREADS : BM_THREADS), smem, stream, P), 0)
  (void)SSD_BEAM_PICK(lm != nullptr, width, SSD_F_GO);
#undef SSD_F_GO
  return check_launch("beam_kernel");
}

}  // namespace ssd
