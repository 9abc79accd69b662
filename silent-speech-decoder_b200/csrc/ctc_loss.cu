// CTC loss, forward + backward in ONE launch: nll[B] and d nll / d log_probs.
//
// Replaces torch.nn.CTCLoss as called at reference src/models/losses.py:56 (ATen ctc_loss + its
// autograd backward).  See DESIGN.md section 4 for the derivation and the roofline discussion.
//
// Numerics.  The alpha/beta recursions run in the LINEAR domain on fp64 values
//   alpha_t[s] = (alpha_{t-1}[s] + alpha_{t-1}[s-1] + skip[s] * alpha_{t-1}[s-2]) * p_t[l'_s]
// (beta: the mirror image, pulling from s+1, s+2).  Both sweeps run THE SAME code: a sweep works in
// "position" space u, with u = s for alpha and u = U - s for beta (U = 32*NC - 2, parity kept), where
// every position pulls from u-1 and u-2.  One instruction stream per role keeps the hot code of a CTA
// inside the SM's instruction cache (the first versions were instruction-fetch bound).  A lane owns
// NC consecutive positions in registers together with ONE integer exponent for them (block floating
// point per lane); every 8 frames the lane renormalises its strip by an exact power of two.  Nothing on the recursion needs
// exp/log (ATen's log-space form costs 3-4 MUFU ops per cell) and there is no global dynamic
// range limit: cells may differ by thousands of binades across the lattice.  p = exp(lp) is
// evaluated once per (frame, label) with an fp32 mantissa and a full fp64 exponent.
//
// Parallel structure: one CTA of 8 warps per utterance (CTAs pull utterances from a queue).
// The two sweeps MEET IN THE MIDDLE.  Per direction d in {forward, backward} ("half"):
//   M  main      runs sweep d over all T frames: first its own half of the frames (phase 1,
//                checkpointing its strip every K frames to an L2-resident workspace), then the
//                other half (phase 2), where it multiplies its fresh values with the OTHER sweep's
//                pre-emission sums to get lattice posteriors and stages them for G
//   H  helper    phase 2 only: re-computes the other sweep K frames at a time from its checkpoints
//                into shared memory (upper 32 bits of each fp64 sum + the per-lane exponents), one
//                segment ahead of M
//   L  loader    issues 1-D TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP) of the next
//                log-prob rows 3 segments ahead and converts them to probabilities for M, H and G
//   G  gradient  one segment behind M: per frame reduces the staged posteriors per label class
//                (blank: shuffle reduce; labels: staging is sorted by class, so one warp prefix
//                scan gives every class sum as a difference) and writes  scale * (p - posterior)
// Alpha and beta are never materialised in HBM: traffic is log_probs once + gradient once +
// the 1/K checkpoint spill.  All hand-offs are mbarrier full/empty pairs; no __syncthreads in
// the frame loop.  Every waiter of a barrier is also part of the chain that lets it complete
// again (or has a barrier of its own), so a parity can never alias.
#include <cmath>
#include <cstring>
#include <cstdlib>

#include "ssd_common.cuh"

namespace ssd {

constexpr int LS_THREADS = 256;
constexpr int LS_NRAW = 4;     // raw-row ring depth per half (TMA prefetch distance = 3 segments)
constexpr int LS_NP = 4;       // probability-row ring depth per half
constexpr int LS_RESCALE = 8;  // frames between strip renormalisations
constexpr int E_ADOPT = 200;   // incoming exponent this far above mine: rebase my strip instead
constexpr int E_NONE = -(1 << 30);

// ------------------------------------------------------------------ scalar helpers
__device__ __forceinline__ int expo_hi(int h) { return ((h >> 20) & 0x7ff) - 1023; }
// 2^e as double, e in [-1022, 1023]
__device__ __forceinline__ double pow2d(int e) { return __hiloint2double((e + 1023) << 20, 0); }
// 2^e, flushing to zero below the normal range; e <= 1023
__device__ __forceinline__ double pow2_flush(int e) { return e < -1022 ? 0.0 : pow2d(e); }
// 2^e for any int e (gradual underflow, overflow to inf)
__device__ __forceinline__ double pow2_wide(int e) {
  if (e < -1070) return 0.0;
  if (e > 1023) return INFINITY;
  if (e >= -1022) return pow2d(e);
  return pow2d(e + 60) * pow2d(-60);
}
// A double cut to its upper 32 bits (sign, 11-bit exponent, 20-bit mantissa, round to nearest):
// the shared-memory format of the helper's sums -- full fp64 range at 4 bytes per cell.
__device__ __forceinline__ int pack_hi(double v) { return __double2hiint(v) + (int)((unsigned)__double2loint(v) >> 31); }
__device__ __forceinline__ double unpack_hi(int h) { return __hiloint2double(h, 0); }
// 2^x for x in [0,1): one MUFU.EX2, no range handling needed
__device__ __forceinline__ float ex2_unit(float x) {
#ifdef SSD_EMU
  return exp2f(x);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
// exp(lp) as a double: fp32 mantissa (one ex2 on the fractional part) but the exponent goes straight
// into the fp64 exponent field, so log-probs down to -700 stay non-zero (fp32 exp dies at -87)
__device__ __forceinline__ double prob_of(float lp) {
  const float t = lp * 1.4426950408889634f;
  if (!(t > -1000.f)) return 0.0;  // also catches -inf and NaN
  const float k = floorf(t);
  const double m = (double)ex2_unit(t - k);  // [1,2]
  return __hiloint2double(__double2hiint(m) + ((int)k << 20), __double2loint(m));
}

struct LossParams {
  const float* lp;
  const float* lse;    // non-null: `lp` holds logits, lse[b * lse_stride + t] their per-frame log-sum-exp (row f-2)
  int64_t lse_stride;
  int64_t T, sb, st;
  int B, V;
  const int64_t* targets;
  int64_t tgt_stride;
  int S_max;
  const int64_t* in_len;
  const int64_t* tg_len;
  const int32_t* order;
  int blank, zero_inf;
  int64_t mean_den;
  const float* grad_mul;
  float* nll;
  float* grad;
  unsigned* counter;
  unsigned char* ckpt;  // per CTA: [2 dirs][nseg_max] strips + exponents
  size_t ckpt_cta_bytes, ckpt_rec_bytes;
  int nseg_max;
  int K;       // frames per segment / pipeline stage
  int NCmax;   // cells per lane the shared-memory stages are sized for
  int VP;      // p-row stride in doubles (V + dummy zero column, even)
  int Spad;    // label staging capacity, multiple of 256
  int GS;      // floats per staged frame: Spad label posteriors | 32 blank partials | flag
  int use_tma;
  long long* tlog;  // timeline log (SSD_LOSS_DBG & 8): [role 4][half 2][512][2] clock64 stamps of CTA 0's first utterance
  int dbg;     // timing experiments only (SSD_LOSS_DBG): 1 G idles, 2 M skips staging, 4 H skips recompute
  // second pass after the fp32 fast path (ctc_loss_fast.cu): only utterances with only[b] != 0 are computed
  const int* only;
  const unsigned* only_count;
  // last resort (rescue_rows below): utterances whose result this kernel could not hold are redone in log space
  int* rescue;            // [B] non-zero = redo
  double* rescue_alpha;   // [T][2 S_max + 1] log alpha of the utterance being redone
};

struct LossSmem {  // byte offsets into dynamic shared memory, identical on host and device
  size_t p_slot, w_slot, we_slot, g_slot, raw_slot;  // sizes
  size_t p_off, w_off, we_off, g_off, raw_off;       // per half: ring bases (half stride = half_bytes)
  size_t half_bytes, lab_off, pos_off, cls_off, bar_off, scal_off, total;
};

// mbarrier indices within a half
constexpr int B_PFULL = 0, B_PEMPTY = LS_NP, B_WFULL = 2 * LS_NP, B_WEMPTY = 2 * LS_NP + 2, B_GFULL = 2 * LS_NP + 4,
              B_GEMPTY = 2 * LS_NP + 6, B_HPFULL = 2 * LS_NP + 8, B_RAW = 3 * LS_NP + 8, B_PER_HALF = 3 * LS_NP + 8 + LS_NRAW;

__host__ __device__ inline LossSmem loss_smem_layout(int K, int NCmax, int V, int VP, int GS, int S_tab) {
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  LossSmem m;
  m.p_slot = al((size_t)K * VP * 8);
  m.w_slot = al((size_t)K * NCmax * 32 * 4);
  m.we_slot = al((size_t)K * 64 * 4);
  m.g_slot = al((size_t)K * GS * 4);
  m.raw_slot = al((size_t)K * V * 4);
  size_t o = 0;
  m.p_off = o; o += LS_NP * m.p_slot;
  m.w_off = o; o += 2 * m.w_slot;
  m.we_off = o; o += 2 * m.we_slot;
  m.g_off = o; o += 2 * m.g_slot;
  m.raw_off = o; o += LS_NRAW * m.raw_slot;
  m.half_bytes = o;
  o = 2 * m.half_bytes;
  m.lab_off = o; o += al((size_t)S_tab * 4);
  m.pos_off = o; o += al((size_t)S_tab * 4);
  m.cls_off = o; o += al((size_t)(V + 2) * 4);
  m.bar_off = o; o += al((size_t)(2 * B_PER_HALF + 2) * 8);
  m.scal_off = o; o += 64;
  m.total = o;
  return m;
}

__device__ __forceinline__ void tstamp(const LossParams& P, int role, int h, int idx, int which, unsigned uiter) {
#ifndef SSD_EMU
  if ((P.dbg & 8) && blockIdx.x == 0 && uiter == 0 && idx < 512 && (threadIdx.x & 31) == 0) P.tlog[((role * 2 + h) * 512 + idx) * 2 + which] = clock64();
#endif
}

// warp-level hand-off: one lane talks to the mbarrier, __syncwarp orders the other lanes' shared-memory
// accesses around it (32x fewer atomics and polls than every lane arriving / waiting)
__device__ __forceinline__ void warp_wait(uint64_t* bar, uint32_t parity) {
  mbar_wait(bar, parity);  // every lane: try_wait parks the whole warp in hardware and wakes it promptly
}
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  __syncwarp();
  if (lane_id() == 0) mbar_arrive(bar);
}

struct LossScal {
  int utt, b, T, S, bad;
  float scale;
};

// ------------------------------------------------------------------ sweep state and steps
// Strip of NC cells per lane (cell = lane*NC + i, even i = blank) with ONE exponent e per lane:
// true value = x[i] * 2^e.  Exponents only move in renorm() + exchange(), every LS_RESCALE frames,
// in lockstep for all lanes; in between, the neighbour's boundary values are brought into my basis
// by the cached factor fin = 2^(neighbour e - my e).  z: my strip was all zero at the last renorm.
template <int NC>
struct Strip {
  double x[NC];
  double fin;
  int e;
  bool z;
};

// this frame's probabilities of my cells, fetched one frame ahead of their use
template <int NC>
struct PRow {
  double lab[NC / 2];
  double blank;
};
template <int NC>
__device__ __forceinline__ void load_prow(PRow<NC>& p, const unsigned char* __restrict__ prow, const int (&lam)[NC / 2], int blank_off) {
  p.blank = *reinterpret_cast<const double*>(prow + blank_off);
#pragma unroll
  for (int k = 0; k < NC / 2; ++k) p.lab[k] = *reinterpret_cast<const double*>(prow + lam[k]);
}

// One frame of the recursion in position space: u pulls from u-1 and (odd u, skip allowed) u-2.
// STORE (helper): also emit the pre-emission sums, packed to 32 bits, at the addresses in wofs
// (the main's layout: position U - u).
template <int NC, bool STORE>
__device__ __forceinline__ void ctc_step(Strip<NC>& s, unsigned skipm, const PRow<NC>& p, int* __restrict__ wdst, int wlast) {
  double (&x)[NC] = s.x;
  double n0 = __shfl_up_sync(FULL, x[NC - 1], 1);
  if (lane_id() == 0) n0 = 0.0;
  n0 *= s.fin;
#pragma unroll
  for (int i = NC - 1; i >= 0; --i) {  // descending: inputs of lower positions are still intact
    const double a1 = (i >= 1) ? x[i >= 1 ? i - 1 : 0] : n0;
    const double a2 = (i >= 2) ? x[i >= 2 ? i - 2 : 0] : n0;
    double sum = x[i] + a1;
    if (i & 1) {
      if ((skipm >> (i >> 1)) & 1u) sum += a2;
      x[i] = sum * p.lab[i >> 1];
    } else {
      x[i] = sum * p.blank;
    }
    if (STORE) {
      // my position lane*NC+i is the main's position U-(lane*NC+i): lane 31-lane, register NC-2-i,
      // except my last register, which lands in register NC-1 of lane 30-lane
      if (i < NC - 1) wdst[(NC - 2 - i) * 32] = pack_hi(sum);
      else if (wlast >= 0) wdst[wlast] = pack_hi(sum);
    }
  }
}

// renormalise my strip so that its largest cell is in [1,2); branch free, tree-shaped maximum
template <int NC>
__device__ __forceinline__ void renorm(Strip<NC>& s) {
  int hw[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) hw[i] = __double2hiint(s.x[i]);
#pragma unroll
  for (int w = 1; w < NC; w <<= 1)
#pragma unroll
    for (int i = 0; i + w < NC; i += 2 * w) hw[i] = hw[i] > hw[i + w] ? hw[i] : hw[i + w];
  const int m = hw[0];
  int ex = expo_hi(m);
  ex = ex < -1022 ? -1022 : ex;
  const bool zero = (m == 0);
  const double f = zero ? 1.0 : pow2d(-ex);
#pragma unroll
  for (int i = 0; i < NC; ++i) s.x[i] *= f;
  s.e += zero ? 0 : ex;
  s.z = zero;
}

// after every renorm: learn the lower neighbour's exponent and cache the factor for its boundary
// value.  An empty strip takes the exponent of the nearest non-empty strip below it, so that the
// first mass reaching it lands close to its basis.
template <int NC>
__device__ __forceinline__ void exchange(Strip<NC>& s) {
  const int lane = lane_id();
  const unsigned live = __ballot_sync(FULL, !s.z);
  const unsigned below = live & ((1u << lane) - 1u);
  const int src = below ? 31 - __clz((int)below) : lane;
  const int e_near = __shfl_sync(FULL, s.e, src);
  if (s.z) s.e = e_near;
  const int ein = __shfl_up_sync(FULL, s.e, 1);
  int d = lane == 0 ? 0 : ein - s.e;
  // an empty neighbour's exponent says nothing about what will arrive: never rebase on it
  const bool nb_live = lane > 0 && ((live >> (lane - 1)) & 1u);
  if (!nb_live) d = d > 1000 ? 1000 : d;
  if (d > E_ADOPT && nb_live) {  // neighbour dwarfs me: rebase my strip (rare, divergent)
    const double f = pow2_wide(-d);
#pragma unroll
    for (int i = 0; i < NC; ++i) s.x[i] *= f;
    s.e = ein;
    d = 0;
  }
  s.fin = pow2_flush(d);
}

// position <-> lattice cell of sweep direction dir:  dir 0: u = s;  dir 1: u = U - s, U = 32*NC - 2
template <int NC>
__device__ __forceinline__ int cell_of(int dir, int u) {
  return dir ? 32 * NC - 2 - u : u;
}

// start state of a sweep: the virtual frame before the first one (the first cell = 1)
template <int NC>
__device__ __forceinline__ void sweep_init(int dir, int L, Strip<NC>& s) {
  const int lane = lane_id();
  const int u0 = dir ? 32 * NC - 2 - (L - 1) : 0;  // position of cell L-1 (backward) or 0 (forward)
#pragma unroll
  for (int i = 0; i < NC; ++i) s.x[i] = (lane * NC + i == u0) ? 1.0 : 0.0;
  s.e = 0;
  s.z = (lane != u0 / NC);
  s.fin = 1.0;
}

// per-lane label column offsets (bytes into a p row), skip mask and (optionally) staging slots for
// the odd positions of my strip, in sweep direction dir
template <int NC>
__device__ __forceinline__ void lane_labels(int dir, int S, int V, const int* __restrict__ lab, const int* __restrict__ pos,
                                            int (&lam)[NC / 2], unsigned& skipm, int* spos) {
  const int lane = lane_id();
  const int L = 2 * S + 1;
  skipm = 0u;
#pragma unroll
  for (int k = 0; k < NC / 2; ++k) {
    const int u = lane * NC + 2 * k + 1;
    const int c = cell_of<NC>(dir, u);  // odd, may be out of the lattice
    int col = V;                        // dummy column (p = 0) outside the target
    int sp = -1;
    if (c >= 1 && c < L) {
      const int j = (c - 1) >> 1;
      col = lab[j];
      if (pos) sp = pos[j];
      const int c2 = cell_of<NC>(dir, u - 2);  // the cell two positions back along the sweep
      if (u >= 2 && c2 >= 1 && c2 < L && lab[(c2 - 1) >> 1] != col) skipm |= 1u << k;
    }
    lam[k] = col * 8;
    if (spos) spos[k] = sp;
  }
}

struct Half {  // shared-memory views of one half's rings
  unsigned char* base;
  uint64_t* bar;
  size_t p_off, p_slot, w_off, w_slot, we_off, we_slot, g_off, g_slot, raw_off, raw_slot;
  __device__ __forceinline__ unsigned char* p(int slot) const { return base + p_off + slot * p_slot; }
  __device__ __forceinline__ int* w(int slot) const { return reinterpret_cast<int*>(base + w_off + slot * w_slot); }
  __device__ __forceinline__ int* we(int slot) const { return reinterpret_cast<int*>(base + we_off + slot * we_slot); }
  __device__ __forceinline__ float* g(int slot) const { return reinterpret_cast<float*>(base + g_off + slot * g_slot); }
  __device__ __forceinline__ float* raw(int slot) const { return reinterpret_cast<float*>(base + raw_off + slot * raw_slot); }
};

// which K-frame block of which sweep direction production q of half h refers to
struct Prod {
  int dir, tau0, n;
};
__device__ __forceinline__ Prod prod_desc(int q, int h, int m_h, int m_o, int K, int NP1, int NSEG) {
  Prod d;
  if (q < NP1) {
    d.dir = h;
    d.tau0 = q * K;
    d.n = min(K, m_h - d.tau0);
  } else {
    const int seg = NSEG - 1 - (q - NP1);
    d.dir = h ^ 1;
    d.tau0 = seg * K;
    d.n = min(K, m_o - d.tau0);
  }
  return d;
}

struct Plan {  // per utterance, per half
  int h, b, Tb, S, L, m_h, m_o, NP1, NSEG, NSEG_RUN, NPROD;
};

// ------------------------------------------------------------------ L: loader / converter
__device__ void loader_role(const LossParams& P, const Half& hf, const Plan& pl, unsigned qbase, unsigned jbase, unsigned uiter) {
  const int lane = lane_id();
  const int K = P.K, V = P.V, VP = P.VP;
  const float* base = P.lp + (int64_t)pl.b * P.sb;
  auto issue_raw = [&](int q) {  // lane 0
    const Prod pd = prod_desc(q, pl.h, pl.m_h, pl.m_o, K, pl.NP1, pl.NSEG);
    const int t_lo = pd.dir ? pl.Tb - pd.tau0 - pd.n : pd.tau0;
    const int slot = (int)((qbase + (unsigned)q) % LS_NRAW);
    const uint32_t bytes = (uint32_t)(pd.n * V * 4);
    mbar_arrive_expect_tx(&hf.bar[B_RAW + slot], bytes);
    bulk_g2s(hf.raw(slot), base + (int64_t)t_lo * P.st, bytes, &hf.bar[B_RAW + slot]);
  };
  if (P.use_tma && lane == 0)
    for (int q = 0; q < pl.NPROD && q < LS_NRAW - 1; ++q) issue_raw(q);
  const int V4 = V >> 2;
  for (int q = 0; q < pl.NPROD; ++q) {
    // every lane has observed raw slot (q-1)'s completion before lane 0 re-arms that slot's mbarrier
    __syncwarp();
    if (P.use_tma && lane == 0 && q + LS_NRAW - 1 < pl.NPROD) issue_raw(q + LS_NRAW - 1);
    const unsigned Q = qbase + (unsigned)q;
    const int slot = (int)(Q % LS_NP);
    warp_wait(&hf.bar[B_PEMPTY + slot], ((Q / LS_NP) & 1u) ^ 1u);
    tstamp(P, 2, pl.h, q, 0, uiter);
    const Prod pd = prod_desc(q, pl.h, pl.m_h, pl.m_o, K, pl.NP1, pl.NSEG);
    double* pR = reinterpret_cast<double*>(hf.p(slot));
    // logits input: lane r holds the log-sum-exp of the block's row r (K <= 32), handed out by shuffle
    float lse_r = 0.f;
    {
      const int t_lo = pd.dir ? pl.Tb - pd.tau0 - pd.n : pd.tau0;
      if (P.lse && lane < pd.n) lse_r = P.lse[(int64_t)pl.b * P.lse_stride + t_lo + lane];
    }
    // row r of the slot <-> local time tau0 + r of direction pd.dir
    if (P.use_tma) {
      const int rs = (int)(Q % LS_NRAW);
      warp_wait(&hf.bar[B_RAW + rs], (Q / LS_NRAW) & 1u);
      const float4* raw4 = reinterpret_cast<const float4*>(hf.raw(rs));
      const int n_it = (pd.n * V4 + 31) / 32;
      for (int it = 0; it < n_it; ++it) {  // V % 4 == 0 on this path
        const int idx0 = it * 32 + lane;
        const bool live = idx0 < pd.n * V4;
        const int idx = live ? idx0 : 0;
        const int br = idx / V4, c4 = idx - br * V4;
        const int r = pd.dir ? pd.n - 1 - br : br;
        const float ls = __shfl_sync(FULL, lse_r, br);
        if (!live) continue;
        float4 v = raw4[idx];
        v.x -= ls; v.y -= ls; v.z -= ls; v.w -= ls;
        double2* dst = reinterpret_cast<double2*>(pR + r * VP + 4 * c4);
        dst[0] = make_double2(prob_of(v.x), prob_of(v.y));
        dst[1] = make_double2(prob_of(v.z), prob_of(v.w));
      }
    } else {
      const int t_lo = pd.dir ? pl.Tb - pd.tau0 - pd.n : pd.tau0;
      for (int r = 0; r < pd.n; ++r) {
        const int br = pd.dir ? pd.n - 1 - r : r;
        const float ls = __shfl_sync(FULL, lse_r, br);
        for (int v = lane; v < V; v += 32) pR[r * VP + v] = prob_of(ld_stream(base + (int64_t)(t_lo + br) * P.st + v) - ls);
      }
    }
    if (lane < pd.n) pR[lane * VP + V] = 0.0;  // dummy column (K <= 32)
    warp_arrive(&hf.bar[B_PFULL + slot]);
    tstamp(P, 2, pl.h, q, 1, uiter);
    if (q >= pl.NP1) warp_arrive(&hf.bar[B_HPFULL + (int)((jbase + (unsigned)(q - pl.NP1)) % LS_NP)]);  // the helper's view
  }
}

// ------------------------------------------------------------------ H: recompute helper (phase 2 only)
template <int NC>
__device__ void helper_role(const LossParams& P, const Half& hf, const Plan& pl, const int* __restrict__ lab, uint64_t* p1done_other,
                            unsigned uiter, unsigned qbase, unsigned jbase, const unsigned char* ckpt_cta) {
  const int lane = lane_id();
  const int K = P.K, VP8 = P.VP * 8, blank_off = P.blank * 8;
  const int dir = pl.h ^ 1;  // I re-compute the OTHER sweep
  int lam[NC / 2];
  unsigned skipm;
  lane_labels<NC>(dir, pl.S, P.V, lab, nullptr, lam, skipm, nullptr);
  // where my sums go in the main's layout (see ctc_step)
  const int wbase = 31 - lane;
  const int wlast = lane < 31 ? (NC - 1) * 32 + (30 - lane) : -1;
  Strip<NC> s;
  for (int j = 0; j < pl.NSEG_RUN; ++j) {
    const unsigned Q = qbase + (unsigned)(pl.NP1 + j), J = jbase + (unsigned)j;
    const int ps = (int)(Q % LS_NP), ws = (int)(J & 1u);
    if (j == 0) warp_wait(p1done_other, uiter & 1u);  // the other main has published its checkpoints
    // rows of production Q are ready.  A barrier of my own (armed for phase-2 productions only):
    // I skip the phase-1 productions, so p_full's parity could alias by the time I get here.
    warp_wait(&hf.bar[B_HPFULL + (int)(J % LS_NP)], (J / LS_NP) & 1u);
    warp_wait(&hf.bar[B_WEMPTY + ws], ((J >> 1) & 1u) ^ 1u);
    tstamp(P, 1, pl.h, pl.NP1 + j, 0, uiter);
    const int seg = pl.NSEG - 1 - j;
    const int tau0 = seg * K, n = min(K, pl.m_o - tau0);
    if (seg == 0) {
      sweep_init<NC>(dir, pl.L, s);
    } else {
      const unsigned char* cp = ckpt_cta + ((size_t)dir * P.nseg_max + seg) * P.ckpt_rec_bytes;
      const double* cx = reinterpret_cast<const double*>(cp);
#pragma unroll
      for (int i = 0; i < NC; ++i) s.x[i] = cx[i * 32 + lane];
      const int* ce = reinterpret_cast<const int*>(cp + (size_t)NC * 32 * 8);
      s.e = ce[lane];
      s.z = ce[32 + lane] != 0;
      s.fin = reinterpret_cast<const double*>(ce + 64)[lane];
    }
    const unsigned char* pR = hf.p(ps);
    int* wR = hf.w(ws);
    int* wE = hf.we(ws);
    int tau = tau0;
    PRow<NC> pc;
    load_prow<NC>(pc, pR, lam, blank_off);
    for (int r = 0; r < n; ++r, ++tau) {
      ctc_step<NC, true>(s, skipm, pc, wR + r * NC * 32 + wbase, wlast - wbase);
      if (r + 1 < n) load_prow<NC>(pc, pR + (r + 1) * VP8, lam, blank_off);  // next frame's, in flight during the rest
      // basis of the sums just stored: exponent A of main lane 31-lane, exponent B of main lane 30-lane
      wE[r * 64 + 2 * (31 - lane)] = s.e;
      if (lane < 31) wE[r * 64 + 2 * (30 - lane) + 1] = s.e;
      if ((tau & (LS_RESCALE - 1)) == LS_RESCALE - 1) {
        renorm<NC>(s);
        exchange<NC>(s);
      }
    }
    warp_arrive(&hf.bar[B_WFULL + ws]);
    tstamp(P, 1, pl.h, pl.NP1 + j, 1, uiter);
  }
}

// ------------------------------------------------------------------ M: main sweep
template <int NC>
__device__ void main_role(const LossParams& P, const Half& hf, const Plan& pl, const int* __restrict__ lab, const int* __restrict__ pos,
                          uint64_t* p1done_mine, unsigned qbase, unsigned jbase, unsigned char* ckpt_cta, unsigned uiter) {
  const int lane = lane_id();
  const int K = P.K, VP8 = P.VP * 8, blank_off = P.blank * 8;
  const int dir = pl.h;
  int lam[NC / 2], spos[NC / 2];
  unsigned skipm;
  lane_labels<NC>(dir, pl.S, P.V, lab, pos, lam, skipm, spos);
  Strip<NC> s;
  sweep_init<NC>(dir, pl.L, s);
  int tau = 0;

  // ---------------- phase 1: my half of the frames, checkpoint every K
  for (int q = 0; q < pl.NP1; ++q) {
    const unsigned Q = qbase + (unsigned)q;
    const int ps = (int)(Q % LS_NP);
    warp_wait(&hf.bar[B_PFULL + ps], (Q / LS_NP) & 1u);
    tstamp(P, 0, pl.h, q, 0, uiter);
    const unsigned char* pR = hf.p(ps);
    const int n = min(K, pl.m_h - q * K);
    PRow<NC> pc;
    load_prow<NC>(pc, pR, lam, blank_off);
    for (int r = 0; r < n; ++r, ++tau) {
      ctc_step<NC, false>(s, skipm, pc, nullptr, 0);
      if (r + 1 < n) load_prow<NC>(pc, pR + (r + 1) * VP8, lam, blank_off);  // next frame's, in flight during the rest
      if ((tau & (LS_RESCALE - 1)) == LS_RESCALE - 1) {
        renorm<NC>(s);
        exchange<NC>(s);
      }
    }
    warp_arrive(&hf.bar[B_PEMPTY + ps]);
    tstamp(P, 0, pl.h, q, 1, uiter);
    if (q + 1 < pl.NP1) {  // state after frame (q+1)K-1 starts segment q+1
      unsigned char* cp = ckpt_cta + ((size_t)dir * P.nseg_max + (q + 1)) * P.ckpt_rec_bytes;
      double* cx = reinterpret_cast<double*>(cp);
#pragma unroll
      for (int i = 0; i < NC; ++i) cx[i * 32 + lane] = s.x[i];
      int* ce = reinterpret_cast<int*>(cp + (size_t)NC * 32 * 8);
      ce[lane] = s.e;
      ce[32 + lane] = s.z ? 1 : 0;
      reinterpret_cast<double*>(ce + 64)[lane] = s.fin;
    }
  }
  __threadfence_block();
  warp_arrive(p1done_mine);

  // ---------------- phase 2: the other half; posteriors = mine (post-emission) x other's (pre-emission)
  bool first = true, feasible = true;
  double invPm = 1.0;
  int eP = 0;
  const bool want_grad = P.grad != nullptr;
  for (int j = 0; j < pl.NSEG_RUN; ++j) {
    const unsigned Q = qbase + (unsigned)(pl.NP1 + j), J = jbase + (unsigned)j;
    const int ps = (int)(Q % LS_NP), ws = (int)(J & 1u);
    const int seg = pl.NSEG - 1 - j;
    const int n = min(K, pl.m_o - seg * K);
    warp_wait(&hf.bar[B_WFULL + ws], (J >> 1) & 1u);
    warp_wait(&hf.bar[B_PFULL + ps], (Q / LS_NP) & 1u);
    if (want_grad) warp_wait(&hf.bar[B_GEMPTY + ws], ((J >> 1) & 1u) ^ 1u);
    tstamp(P, 0, pl.h, pl.NP1 + j, 0, uiter);
    const unsigned char* pR = hf.p(ps);
    const int* wR = hf.w(ws) + lane;
    const int* wE = hf.we(ws) + 2 * lane;
    float* gR = hf.g(ws);
    PRow<NC> pc;
    load_prow<NC>(pc, pR + (n - 1) * VP8, lam, blank_off);
    for (int r = n - 1; r >= 0; --r, ++tau) {
      // the other sweep's sums for my positions (issued before the step so their latency is covered)
      const int* wo = wR + r * NC * 32;
      int wv[NC];
#pragma unroll
      for (int i = 0; i < NC; ++i) wv[i] = wo[i * 32];
      if (lane == 31) wv[NC - 1] = 0;  // position 32*NC-1 belongs to no cell
      const int e_a = wE[r * 64], e_b = wE[r * 64 + 1];  // bases of registers 0..NC-2 / NC-1
      ctc_step<NC, false>(s, skipm, pc, nullptr, 0);
      if (r > 0) load_prow<NC>(pc, pR + (r - 1) * VP8, lam, blank_off);  // next frame's, in flight during the staging
      const int q_l = s.e + e_a;  // exponent of my lane's products
      int db = e_b - e_a;
      db = db > 1000 ? 1000 : db;
      const double f_b = lane == 31 ? 0.0 : pow2_flush(db);  // brings register NC-1's product into that basis
      if (first) {
        first = false;
        // P = sum over the lattice, combined across lanes at the largest lane exponent
        double sl = 0.0;
#pragma unroll
        for (int i = 0; i < NC; ++i) sl += s.x[i] * unpack_hi(wv[i]) * (i == NC - 1 ? f_b : 1.0);
        const bool ok = sl > 0.0 && sl < INFINITY;
        const int key = ok ? q_l + expo_hi(__double2hiint(sl)) : E_NONE;
        const int Qm = __reduce_max_sync(FULL, key);
        feasible = Qm > E_NONE;
        if (feasible) {
          double tot = ok ? sl * pow2_wide(q_l - Qm) : 0.0;
#pragma unroll
          for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
          const int et = expo_hi(__double2hiint(tot));
          const double pm = tot * pow2_wide(-et);  // in [1,2)
          invPm = 1.0 / pm;
          eP = Qm + et;
          if (dir == 0 && lane == 0) P.nll[pl.b] = (float)(-(log(pm) + (double)eP * 0.69314718055994530942));
        } else if (dir == 0 && lane == 0) {
          P.nll[pl.b] = P.zero_inf ? 0.f : INFINITY;
          // infeasible, or lost?  A target fits iff its labels plus one blank per repeated label fit the frames;
          // if it does, the lattice sum was beyond this kernel's reach (thousands of binades inside one strip)
          int need = pl.S;
          for (int j = 1; j < pl.S; ++j) need += lab[j] == lab[j - 1] ? 1 : 0;
          if (need <= pl.Tb) atomicOr(&P.rescue[pl.b], 1);
        }
        if (!want_grad) break;
      }
      // stage this frame for G: label posteriors in class-sorted order + my lane's blank partial
      float* gf = gR + (size_t)r * P.GS;
      int kq = q_l - eP;
      kq = kq > 1000 ? 1000 : kq;
      const double c = feasible ? pow2_flush(kq) * invPm : 0.0;
      double gb = 0.0;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        const double pr = s.x[i] * unpack_hi(wv[i]);
        if (i & 1) {
          if (spos[i >> 1] >= 0) gf[spos[i >> 1]] = (float)(pr * (i == NC - 1 ? c * f_b : c));
        } else {
          gb += pr;
        }
      }
      gf[P.Spad + lane] = (float)(gb * c);
      if ((tau & (LS_RESCALE - 1)) == LS_RESCALE - 1) {
        renorm<NC>(s);
        exchange<NC>(s);
      }
    }
    if (want_grad) {
      if (lane == 0) gR[P.Spad + 32] = feasible ? 1.f : 0.f;  // segment flag (frame 0's spare float)
      warp_arrive(&hf.bar[B_GFULL + ws]);
    }
    warp_arrive(&hf.bar[B_WEMPTY + ws]);
    warp_arrive(&hf.bar[B_PEMPTY + ps]);
    tstamp(P, 0, pl.h, pl.NP1 + j, 1, uiter);
  }
}

// ------------------------------------------------------------------ G: gradient writer
// Takes part in every production's release (p_empty) so that the loader can never lap it.
__device__ void grad_role(const LossParams& P, const Half& hf, const Plan& pl, const int* __restrict__ cls, float scale, unsigned qbase,
                          unsigned jbase, unsigned uiter) {
  const int lane = lane_id();
  const int K = P.K, V = P.V, S = pl.S, h = pl.h, VP = P.VP;
  float* grow_base = P.grad + (int64_t)pl.b * P.sb;
  for (int q = 0; q < pl.NP1; ++q) {  // phase 1: nothing to write yet, just keep in step
    const unsigned Q = qbase + (unsigned)q;
    const int ps = (int)(Q % LS_NP);
    warp_wait(&hf.bar[B_PFULL + ps], (Q / LS_NP) & 1u);
    warp_arrive(&hf.bar[B_PEMPTY + ps]);
  }
  int tau = pl.m_h;
  for (int j = 0; j < pl.NSEG_RUN; ++j) {
    const unsigned Q = qbase + (unsigned)(pl.NP1 + j), J = jbase + (unsigned)j;
    const int ps = (int)(Q % LS_NP), ws = (int)(J & 1u);
    const int seg = pl.NSEG - 1 - j;
    const int n = min(K, pl.m_o - seg * K);
    warp_wait(&hf.bar[B_GFULL + ws], (J >> 1) & 1u);  // implies the rows of production Q are in place
    tstamp(P, 3, pl.h, pl.NP1 + j, 0, uiter);
    float* gR = hf.g(ws);
    const double* pR = reinterpret_cast<const double*>(hf.p(ps));
    const bool feasible = gR[P.Spad + 32] != 0.f;
    // FR frames at a time: their scans, reductions and loads are independent chains that the
    // scheduler can interleave, which is what hides the shuffle / shared-memory latencies here
    constexpr int FR = 3;
    for (int r = n - 1; r >= 0; r -= FR) {
      const int nf = r + 1 < FR ? r + 1 : FR;  // frames r, r-1, .., r-nf+1
      const int tbase = tau + (n - 1 - r);     // local time of frame r
      float* gfp[FR];
      float* growp[FR];
#pragma unroll
      for (int f = 0; f < FR; ++f) {
        const int rr = f < nf ? r - f : r;
        gfp[f] = gR + (size_t)rr * P.GS;
        const int tt = tbase + (f < nf ? f : 0);
        growp[f] = grow_base + (int64_t)(h ? pl.Tb - 1 - tt : tt) * P.st;
      }
      if (feasible) {
        float gb[FR], carry[FR];
#pragma unroll
        for (int f = 0; f < FR; ++f) {
          gb[f] = gfp[f][P.Spad + lane];
          carry[f] = 0.f;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1)
#pragma unroll
          for (int f = 0; f < FR; ++f) gb[f] += __shfl_xor_sync(FULL, gb[f], o);
        // inclusive prefix sum over the class-sorted label posteriors, in place, 256 per pass
        for (int base = 0; base < S; base += 256) {
          const int i0 = base + lane * 8;
          float4 a[FR], b[FR];
          float incl[FR], tot8[FR];
#pragma unroll
          for (int f = 0; f < FR; ++f) {
            const float4* p4 = reinterpret_cast<const float4*>(gfp[f] + i0);
            a[f] = p4[0];
            b[f] = p4[1];
          }
#pragma unroll
          for (int f = 0; f < FR; ++f) {
            float4 &A = a[f], &B = b[f];
            A.x = i0 + 0 < S ? A.x : 0.f; A.y = i0 + 1 < S ? A.y : 0.f; A.z = i0 + 2 < S ? A.z : 0.f; A.w = i0 + 3 < S ? A.w : 0.f;
            B.x = i0 + 4 < S ? B.x : 0.f; B.y = i0 + 5 < S ? B.y : 0.f; B.z = i0 + 6 < S ? B.z : 0.f; B.w = i0 + 7 < S ? B.w : 0.f;
            A.y += A.x; A.z += A.y; A.w += A.z; B.x += A.w; B.y += B.x; B.z += B.y; B.w += B.z;
            incl[f] = tot8[f] = B.w;
          }
#pragma unroll
          for (int o = 1; o < 32; o <<= 1)
#pragma unroll
            for (int f = 0; f < FR; ++f) {
              const float y = __shfl_up_sync(FULL, incl[f], o);
              if (lane >= o) incl[f] += y;
            }
#pragma unroll
          for (int f = 0; f < FR; ++f) {
            const float off = carry[f] + incl[f] - tot8[f];  // everything before my 8
            float4 &A = a[f], &B = b[f];
            A.x += off; A.y += off; A.z += off; A.w += off; B.x += off; B.y += off; B.z += off; B.w += off;
            if (f < nf) {
              float4* p4 = reinterpret_cast<float4*>(gfp[f] + i0);
              p4[0] = A;
              p4[1] = B;
            }
            carry[f] += __shfl_sync(FULL, incl[f], 31);
          }
        }
        __syncwarp();
        float mass[FR];
#pragma unroll
        for (int f = 0; f < FR; ++f) mass[f] = 0.f;
        for (int v = lane; v < V; v += 32) {
          const int ca = cls[v], cb = cls[v + 1];
          float gv[FR], pv[FR];
#pragma unroll
          for (int f = 0; f < FR; ++f) {
            gv[f] = cb > ca ? gfp[f][cb - 1] - (ca > 0 ? gfp[f][ca - 1] : 0.f) : 0.f;
            pv[f] = (float)pR[(f < nf ? r - f : r) * VP + v];
          }
#pragma unroll
          for (int f = 0; f < FR; ++f) {
            if (v == P.blank) gv[f] += gb[f];
            mass[f] += gv[f];
            if (f < nf) st_stream(growp[f] + v, (pv[f] - gv[f]) * scale);
          }
        }
        // the posteriors of a frame sum to one; if they do not (or are not numbers) the lattice was beyond what one
        // exponent per strip can hold, and the utterance is redone in log space after the launch's last utterance
#pragma unroll
        for (int o = 16; o; o >>= 1)
#pragma unroll
          for (int f = 0; f < FR; ++f) mass[f] += __shfl_xor_sync(FULL, mass[f], o);
        {
          bool off = false;
#pragma unroll
          for (int f = 0; f < FR; ++f) off = off || (f < nf && !(fabsf(mass[f] - 1.f) <= 1e-4f));
          if (off && lane == 0) atomicOr(&P.rescue[pl.b], 2);
        }
        __syncwarp();
      } else {
        const float fill = P.zero_inf ? 0.f : NAN;
#pragma unroll
        for (int f = 0; f < FR; ++f)
          if (f < nf)
            for (int v = lane; v < V; v += 32) st_stream(growp[f] + v, fill);
      }
    }
    tau += n;
    warp_arrive(&hf.bar[B_GEMPTY + ws]);
    warp_arrive(&hf.bar[B_PEMPTY + ps]);
    tstamp(P, 3, pl.h, pl.NP1 + j, 1, uiter);
  }
}

// ------------------------------------------------------------------ last resort: one utterance in log space
// Plain log-space alpha/beta in fp64 (Graves 2006 eqs. 6-8, 10-11, 16 with ATen's conventions), alpha kept in global
// memory, one CTA, one barrier per frame.  Slow (milliseconds per utterance) and always right: no exponent sharing, no
// range to leave.  Run by the launch's LAST CTA for the utterances the kernel above flagged: posteriors that did not
// sum to one, or a feasible target whose lattice sum came out empty -- confident posteriors forced onto an alignment
// that contradicts them frame after frame put thousands of binades between the cells of ONE strip, which one exponent
// per strip cannot hold (round-2 fuzz: 164 labels on 175 frames of a posterior laid out for 333; nll 1155).
__device__ __forceinline__ double lse3(double a, double b, double c) {
  const double m = fmax(a, fmax(b, c));
  if (m == -INFINITY) return m;
  return m + log(exp(a - m) + exp(b - m) + exp(c - m));
}

__device__ void rescue_row(const LossParams& P, int b, int* lab, double* rowbuf) {
  const int tid = (int)threadIdx.x;
  const int64_t tl64 = P.in_len[b], sl64 = P.tg_len[b];
  const int Tb = (int)tl64, S = (int)sl64, L = 2 * S + 1, V = P.V;
  const float* x = P.lp + (int64_t)b * P.sb;
  const float* lse = P.lse ? P.lse + (int64_t)b * P.lse_stride : nullptr;
  double* alpha = P.rescue_alpha;  // [Tb][L]
  for (int j = tid; j < S; j += LS_THREADS) lab[j] = (int)P.targets[(int64_t)b * P.tgt_stride + j];
  __syncthreads();
  auto lp = [&](int t, int v) { return (double)(x[(int64_t)t * P.st + v] - (lse ? lse[t] : 0.f)); };
  auto cls_of = [&](int s_) { return (s_ & 1) ? lab[s_ >> 1] : P.blank; };
  auto skip_ok = [&](int s_) { return (s_ & 1) && s_ >= 3 && lab[s_ >> 1] != lab[(s_ >> 1) - 1]; };
  // forward
  for (int s_ = tid; s_ < L; s_ += LS_THREADS) alpha[s_] = s_ < 2 ? lp(0, cls_of(s_)) : -INFINITY;
  __syncthreads();
  for (int t = 1; t < Tb; ++t) {
    const double* ap = alpha + (size_t)(t - 1) * L;
    double* ac = alpha + (size_t)t * L;
    for (int s_ = tid; s_ < L; s_ += LS_THREADS)
      ac[s_] = lse3(ap[s_], s_ >= 1 ? ap[s_ - 1] : -INFINITY, skip_ok(s_) ? ap[s_ - 2] : -INFINITY) + lp(t, cls_of(s_));
    __syncthreads();
  }
  const double* aT = alpha + (size_t)(Tb - 1) * L;
  const double logZ = lse3(aT[L - 1], L >= 2 ? aT[L - 2] : -INFINITY, -INFINITY);
  const bool feasible = logZ > -INFINITY;
  if (tid == 0) P.nll[b] = feasible ? (float)(-logZ) : (P.zero_inf ? 0.f : INFINITY);
  if (!P.grad) return;
  float mul = P.grad_mul ? *P.grad_mul : 1.f;
  if (P.mean_den > 0) mul /= (float)((double)P.mean_den * (double)(S > 1 ? S : 1));
  float* g = P.grad + (int64_t)b * P.sb;
  if (!feasible) {
    const float fill = P.zero_inf ? 0.f : NAN;
    for (int i = tid; i < Tb * V; i += LS_THREADS) g[(int64_t)(i / V) * P.st + (i % V)] = fill;
    return;
  }
  // backward: beta rows alternate between the two halves of rowbuf; cab = alpha + beta of the current frame
  double* beta[2] = {rowbuf, rowbuf + L};
  double* cab = rowbuf + 2 * L;
  for (int t = Tb - 1; t >= 0; --t) {
    double* bc = beta[t & 1];
    const double* bn = beta[(t + 1) & 1];
    const double* at = alpha + (size_t)t * L;
    for (int s_ = tid; s_ < L; s_ += LS_THREADS) {
      double v;
      if (t == Tb - 1) {
        v = s_ >= L - 2 ? lp(t, cls_of(s_)) : -INFINITY;
      } else {
        const bool skip_fwd = (s_ & 1) && s_ + 2 < L && lab[s_ >> 1] != lab[(s_ >> 1) + 1];
        v = lse3(bn[s_], s_ + 1 < L ? bn[s_ + 1] : -INFINITY, skip_fwd ? bn[s_ + 2] : -INFINITY) + lp(t, cls_of(s_));
      }
      bc[s_] = v;
      cab[s_] = at[s_] + v;
    }
    __syncthreads();
    for (int v = tid; v < V; v += LS_THREADS) {
      // log sum over the cells of class v, two passes (maximum, then the sum)
      double m = -INFINITY;
      for (int s_ = 0; s_ < L; ++s_)
        if (cls_of(s_) == v) m = fmax(m, cab[s_]);
      double acc = 0.0;
      if (m > -INFINITY)
        for (int s_ = 0; s_ < L; ++s_)
          if (cls_of(s_) == v) acc += exp(cab[s_] - m);
      const double l = lp(t, v);
      const double post = m > -INFINITY ? exp(m + log(acc) - logZ - l) : 0.0;
      g[(int64_t)t * P.st + v] = (float)((exp(l) - post) * (double)mul);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ kernel
template <int NC, int MINB>
__global__ void __launch_bounds__(LS_THREADS, MINB) loss_kernel(const LossParams P) {
  SSD_DYN_SMEM(smem);
  const LossSmem M = loss_smem_layout(P.K, P.NCmax, P.V, P.VP, P.GS, P.S_max + 4);
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // warp -> (role, half).  Scheduler partition = warp % 4: M_0+G_0 | M_1+G_1 | H_0+L_0 | H_1+L_1, so in
  // phase 1 (only M and L busy) the mains do not share an issue port with the loaders
  const int h = warp & 1, role = warp >> 1;  // 0 M, 1 H, 2 G, 3 L

  int* lab = reinterpret_cast<int*>(smem + M.lab_off);
  int* pos = reinterpret_cast<int*>(smem + M.pos_off);
  int* cls = reinterpret_cast<int*>(smem + M.cls_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + M.bar_off);
  LossScal* sc = reinterpret_cast<LossScal*>(smem + M.scal_off);
  uint64_t* p1done = bars + 2 * B_PER_HALF;

  Half hf;
  hf.base = smem + (size_t)h * M.half_bytes;
  hf.bar = bars + h * B_PER_HALF;
  hf.p_off = M.p_off; hf.p_slot = M.p_slot; hf.w_off = M.w_off; hf.w_slot = M.w_slot; hf.we_off = M.we_off; hf.we_slot = M.we_slot;
  hf.g_off = M.g_off; hf.g_slot = M.g_slot; hf.raw_off = M.raw_off; hf.raw_slot = M.raw_slot;

  if (tid == 0) {
    for (int hh = 0; hh < 2; ++hh) {
      uint64_t* b = bars + hh * B_PER_HALF;
      for (int i = 0; i < B_RAW; ++i) mbar_init(&b[i], 1);  // one arrival per arriving warp
      if (P.grad)
        for (int i = 0; i < LS_NP; ++i) mbar_init(&b[B_PEMPTY + i], 2);  // rows are released by M and G
      for (int i = 0; i < LS_NRAW; ++i) mbar_init(&b[B_RAW + i], 1);  // lane 0 + tx bytes
    }
    mbar_init(&p1done[0], 1);
    mbar_init(&p1done[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  unsigned char* ckpt_cta = P.ckpt + (size_t)blockIdx.x * P.ckpt_cta_bytes;
  unsigned qbase = 0u, jbase = 0u, uiter = 0u;
  if (P.only && *P.only_count == 0u) return;  // nothing was flagged by the fast path

  for (;; ++uiter) {
    // ---------------- next utterance + its label tables
    if (tid == 0) {
      int u = (int)atomicAdd(P.counter, 1u);
      if (P.only)
        while (u < P.B && P.only[P.order ? P.order[u] : u] == 0) u = (int)atomicAdd(P.counter, 1u);
      sc->utt = u;
      if (u < P.B) {
        const int b = P.order ? P.order[u] : u;
        const int64_t tl = P.in_len[b], sl = P.tg_len[b];
        sc->b = b;
        sc->bad = (tl < 0 || tl > P.T || sl < 0 || sl > P.S_max) ? 1 : 0;
        sc->T = (int)(tl < 0 ? 0 : (tl > P.T ? P.T : tl));
        sc->S = sc->bad ? 0 : (int)sl;
        float mul = P.grad_mul ? *P.grad_mul : 1.f;
        if (P.mean_den > 0) mul /= (float)((double)P.mean_den * (double)(sl > 1 ? sl : 1));
        sc->scale = mul;
      }
    }
    __syncthreads();
    if (sc->utt >= P.B) break;
    const int b = sc->b, S = sc->S;
    for (int j = tid; j < S; j += LS_THREADS) {
      const int64_t y = P.targets[(int64_t)b * P.tgt_stride + j];
      if (y < 0 || y >= P.V) sc->bad = 1;  // benign race: every writer stores 1
      lab[j] = (int)(y < 0 ? 0 : (y >= P.V ? P.V - 1 : y));
    }
    __syncthreads();
    // class histogram -> start offsets; position of every label in class-sorted order (stable)
    for (int v = tid; v < P.V; v += LS_THREADS) {
      int c = 0;
      for (int j = 0; j < S; ++j) c += (lab[j] == v) ? 1 : 0;
      cls[v + 1] = c;
    }
    __syncthreads();
    if (tid == 0) {
      cls[0] = 0;
      for (int v = 0; v < P.V; ++v) cls[v + 1] += cls[v];
    }
    __syncthreads();
    for (int j = tid; j < S; j += LS_THREADS) {
      const int y = lab[j];
      int k = 0;
      for (int i = 0; i < j; ++i) k += (lab[i] == y) ? 1 : 0;
      pos[j] = cls[y] + k;
    }
    __syncthreads();

    const int bad = sc->bad;
    const int Tb = bad ? 0 : sc->T;
    const float scale = sc->scale;
    // rows the sweeps never visit: frames beyond the input length (all frames if the row is unusable)
    if (P.grad) {
      float* g0 = P.grad + (int64_t)b * P.sb;
      const float fill = bad ? NAN : 0.f;
      for (int64_t r = Tb + warp; r < P.T; r += LS_THREADS / 32)
        for (int v = lane; v < P.V; v += 32) st_stream(g0 + r * P.st + v, fill);
    }
    if (Tb == 0 && tid == 0) P.nll[b] = bad ? NAN : (S == 0 ? 0.f : (P.zero_inf ? 0.f : INFINITY));

    Plan pl;
    pl.h = h; pl.b = b; pl.Tb = Tb; pl.S = S; pl.L = 2 * S + 1;
    const int m = Tb / 2;
    pl.m_h = h == 0 ? m : Tb - m;
    pl.m_o = Tb - pl.m_h;
    pl.NP1 = (pl.m_h + P.K - 1) / P.K;
    pl.NSEG = (pl.m_o + P.K - 1) / P.K;
    // forward only: just the first phase-2 frame of the forward half is needed (it yields P)
    pl.NSEG_RUN = P.grad ? pl.NSEG : (h == 0 ? min(pl.NSEG, 1) : 0);
    pl.NPROD = pl.NP1 + pl.NSEG_RUN;
    if (role == 3) {
      loader_role(P, hf, pl, qbase, jbase, uiter);
    } else if (role == 2) {
      if (P.grad) grad_role(P, hf, pl, cls, scale, qbase, jbase, uiter);
    } else if (role == 1) {
      helper_role<NC>(P, hf, pl, lab, &p1done[h ^ 1], uiter, qbase, jbase, ckpt_cta);  // re-computes direction h^1
    } else {
      main_role<NC>(P, hf, pl, lab, pos, &p1done[h], qbase, jbase, ckpt_cta, uiter);  // sweeps direction h
    }
    qbase += (unsigned)pl.NPROD;
    jbase += (unsigned)pl.NSEG_RUN;
    __syncthreads();  // label tables are rebuilt for the next utterance
  }
  // ---------------- the last CTA to get here redoes what was flagged for the log-space path (normally nothing)
  __shared__ int is_last;
  if (tid == 0) {
    __threadfence();
    is_last = atomicAdd(P.counter + 2, 1u) == gridDim.x - 1u ? 1 : 0;
  }
  __syncthreads();
  if (!is_last || !P.rescue_alpha) return;
  __threadfence();
  for (int bb = 0; bb < P.B; ++bb) {
    if (*reinterpret_cast<volatile int*>(P.rescue + bb) == 0) continue;  // uniform: everybody reads the same word
    // lab (shared) is free now; the beta rows and the per-frame products live at the start of the dynamic shared memory
    rescue_row(P, bb, lab, reinterpret_cast<double*>(smem));
    __syncthreads();
  }
}

// ------------------------------------------------------------------ host side
struct LossPlan {
  int K, NCmax, VP, Spad, GS, ctas_per_sm, grid, nseg_max;
  size_t smem, ckpt_cta_bytes, ckpt_rec_bytes, ws_bytes;
  size_t head_bytes, off_ckpt, off_tlog, off_rescue_alpha;  // workspace: [counters | rescue flags][checkpoints][timeline][log alpha]
  bool rescue_fits;  // the log-space path's three rows fit the shared memory the launch has anyway
};

static int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s && *s ? std::atoi(s) : dflt;
}

static int device_sms() {
#ifdef SSD_EMU
  return 2;
#else
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
#endif
}

static bool plan_for(int64_t B, int64_t T, int64_t V, int64_t S_max, LossPlan& pl) {
  // one code variant per launch, chosen from the widest possible lattice: 2*S_max+1 cells plus the
  // unused top position must fit 32*NC positions
  const int need = (int)((2 * S_max + 2 + 31) / 32);
  const int buckets[5] = {4, 8, 14, 20, 32};
  pl.NCmax = 0;
  for (int i = 0; i < 5 && !pl.NCmax; ++i)
    if (need <= buckets[i]) pl.NCmax = buckets[i];
  if (!pl.NCmax) return false;
  pl.VP = (int)((V + 2) & ~1LL);
  pl.Spad = (int)((S_max + 255) / 256 * 256);
  if (pl.Spad == 0) pl.Spad = 256;
  pl.GS = pl.Spad + 32 + 4;
  const size_t cap = 227 * 1024;
  auto bytes = [&](int K) { return loss_smem_layout(K, pl.NCmax, (int)V, pl.VP, pl.GS, (int)S_max + 4).total; };
  const int kforce = env_int("SSD_LOSS_K", 0);
  int ctas = env_int("SSD_LOSS_CTAS", 0);
  const int max_ctas = pl.NCmax <= 14 ? 2 : 1;  // register budget of the two kernel variants
  int K = 0;
  if (kforce > 0) {
    K = kforce > 32 ? 32 : kforce;
    if (bytes(K) > cap) return false;
    if (ctas <= 0) ctas = (int)(cap / (bytes(K) + 1024));
  } else {
    // prefer two resident CTAs per SM (16 warps) when that still allows K >= 4
    const int want = ctas > 0 ? ctas : max_ctas;
    for (int c = want < max_ctas ? want : max_ctas; c >= 1 && K == 0; --c) {
      int k = 16;
      while (k >= 2 && (bytes(k) + 1024) * c > cap) --k;
      if (k >= (c > 1 ? 4 : 2)) {
        K = k;
        ctas = c;
      }
    }
    if (K == 0) return false;
  }
  if (ctas < 1) ctas = 1;
  if (ctas > max_ctas) ctas = max_ctas;
  pl.K = K;
  pl.ctas_per_sm = ctas;
  pl.smem = bytes(K);
  const int64_t slots = (int64_t)device_sms() * ctas;
  pl.grid = (int)(B < slots ? B : slots);
  pl.nseg_max = (int)((T / 2 + 1 + K - 1) / K + 2);
  pl.ckpt_rec_bytes = (size_t)pl.NCmax * 32 * 8 + 512;  // strip + per-lane e, z, fin
  pl.ckpt_cta_bytes = (size_t)2 * pl.nseg_max * pl.ckpt_rec_bytes;
  pl.head_bytes = 256 + (((size_t)B * 4 + 255) & ~(size_t)255);
  pl.off_ckpt = pl.head_bytes;
  pl.off_tlog = pl.off_ckpt + (size_t)pl.grid * pl.ckpt_cta_bytes;
  pl.off_rescue_alpha = (pl.off_tlog + 8 * 512 * 2 * 8 + 255) & ~(size_t)255;
  pl.ws_bytes = pl.off_rescue_alpha + (size_t)T * (size_t)(2 * S_max + 1) * 8;
  pl.rescue_fits = (size_t)3 * (size_t)(2 * S_max + 1) * 8 <= 2 * loss_smem_layout(pl.K, pl.NCmax, (int)V, pl.VP, pl.GS, (int)S_max + 4).half_bytes;
  return true;
}

// fp32 fast path (ctc_loss_fast.cu)
size_t loss_fast_workspace_bytes(int64_t B, int64_t T, int64_t V, int64_t S_max);
int loss_fast_launch(const float* lp, const float* lse, int64_t lse_stride, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st,
                     const int64_t* targets, int64_t tgt_stride, int64_t S_max, const int64_t* in_len, const int64_t* tg_len,
                     const int32_t* order, int blank, int zero_inf, int64_t mean_den, const float* grad_mul, float* nll, float* grad,
                     void* ws, size_t ws_bytes, const int** flags, const unsigned** nflag, cudaStream_t stream);

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

// workspace = [exact path | fast path]
size_t loss_workspace_bytes(int64_t B, int64_t T, int64_t V, int64_t S_max) {
  LossPlan pl;
  const size_t exact = plan_for(B, T, V, S_max, pl) ? pl.ws_bytes : 256;
  return align256(exact) + loss_fast_workspace_bytes(B, T, V, S_max);
}

size_t loss_fallback_counter_offset(int64_t B, int64_t T, int64_t V, int64_t S_max) {
  if (!loss_fast_workspace_bytes(B, T, V, S_max)) return (size_t)-1;
  LossPlan pl;
  const size_t exact = plan_for(B, T, V, S_max, pl) ? pl.ws_bytes : 256;
  return align256(exact) + 4;  // FastParams::counter[1]
}

template <int NC, int MINB>
static int loss_launch_t(LossParams& P, const LossPlan& pl, cudaStream_t stream) {
  auto kfn = loss_kernel<NC, MINB>;
#ifndef SSD_EMU
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "loss: cudaFuncSetAttribute(%zu): %s", pl.smem, cudaGetErrorString(e));
  e = cudaMemsetAsync(P.counter, 0, pl.head_bytes, stream);  // work queue, exit counter, rescue flags
  if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "loss: cudaMemsetAsync: %s", cudaGetErrorString(e));
#else
  std::memset(P.counter, 0, pl.head_bytes);
#endif
  SSD_LAUNCH(kfn, (unsigned)pl.grid, LS_THREADS, pl.smem, stream, P);
  return check_launch("loss_kernel");
}

int loss_launch(const float* lp, const float* lse, int64_t lse_stride, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st,
                const int64_t* targets, int64_t tgt_stride, int64_t S_max, const int64_t* in_len, const int64_t* tg_len,
                const int32_t* order, int blank, int zero_inf, int64_t mean_den, const float* grad_mul, float* nll, float* grad,
                void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (B == 0) return SSD_OK;
  LossPlan pl;
  if (!plan_for(B, T, V, S_max, pl))
    return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_loss: no shared-memory plan for V=%lld S_max=%lld", (long long)V, (long long)S_max);
  if (!ws || ws_bytes < pl.ws_bytes) return fail(SSD_ERR_WORKSPACE, "ssd_ctc_loss: workspace %zu bytes, need %zu", ws_bytes, pl.ws_bytes);
  // fp32 fast path first; whatever it flags (and everything it does not cover) is done by the fp64 kernel below
  const int* only = nullptr;
  const unsigned* only_count = nullptr;
  {
    const size_t fast_off = align256(pl.ws_bytes);
    const size_t fast_need = loss_fast_workspace_bytes(B, T, V, S_max);
    if (fast_need) {
      if (ws_bytes < fast_off + fast_need)
        return fail(SSD_ERR_WORKSPACE, "ssd_ctc_loss: workspace %zu bytes, need %zu", ws_bytes, fast_off + fast_need);
      const int rc = loss_fast_launch(lp, lse, lse_stride, B, T, V, sb, st, targets, tgt_stride, S_max, in_len, tg_len, order, blank, zero_inf, mean_den,
                                      grad_mul, nll, grad, reinterpret_cast<unsigned char*>(ws) + fast_off, fast_need, &only, &only_count,
                                      stream);
      if (rc != SSD_OK && rc != SSD_ERR_UNSUPPORTED) return rc;
      if (rc == SSD_ERR_UNSUPPORTED) only = nullptr, only_count = nullptr;
    }
  }
  if (only && env_int("SSD_LOSS_FASTONLY", 0)) return SSD_OK;  // timing experiments only: flagged rows stay unrecomputed
  LossParams P;
  P.only = only; P.only_count = only_count;
  P.lp = lp; P.lse = lse; P.lse_stride = lse_stride; P.T = T; P.sb = sb; P.st = st; P.B = (int)B; P.V = (int)V;
  P.targets = targets; P.tgt_stride = tgt_stride; P.S_max = (int)S_max;
  P.in_len = in_len; P.tg_len = tg_len; P.order = order;
  P.blank = blank; P.zero_inf = zero_inf; P.mean_den = mean_den; P.grad_mul = grad_mul;
  P.nll = nll; P.grad = grad;
  P.counter = reinterpret_cast<unsigned*>(ws);
  P.rescue = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(ws) + 256);
  P.rescue_alpha = pl.rescue_fits && !env_int("SSD_LOSS_NORESCUE", 0) ? reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(ws) + pl.off_rescue_alpha) : nullptr;
  P.ckpt = reinterpret_cast<unsigned char*>(ws) + pl.off_ckpt;
  P.ckpt_cta_bytes = pl.ckpt_cta_bytes; P.ckpt_rec_bytes = pl.ckpt_rec_bytes; P.nseg_max = pl.nseg_max;
  P.K = pl.K; P.NCmax = pl.NCmax; P.VP = pl.VP; P.Spad = pl.Spad; P.GS = pl.GS;
  P.use_tma = (st == V) && aligned16(lp) && ((sb * 4) % 16 == 0) && ((V * 4) % 16 == 0) ? 1 : 0;
  if (env_int("SSD_LOSS_NOTMA", 0)) P.use_tma = 0;  // tuning knob
  P.dbg = env_int("SSD_LOSS_DBG", 0);
  P.tlog = reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(ws) + pl.off_tlog);
  switch (pl.NCmax) {
    case 4: return loss_launch_t<4, 2>(P, pl, stream);
    case 8: return loss_launch_t<8, 2>(P, pl, stream);
    case 14: return loss_launch_t<14, 2>(P, pl, stream);
    case 20: return loss_launch_t<20, 1>(P, pl, stream);
    default: return loss_launch_t<32, 1>(P, pl, stream);
  }
}

}  // namespace ssd
