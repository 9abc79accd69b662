// CTC loss, forward + backward: the fp32 FAST PATH (nll[B] and d nll / d log_probs in one launch).
//
// Replaces torch.nn.CTCLoss as called at reference src/models/losses.py:56 together with ctc_loss.cu,
// which stays the exact fp64 path: this kernel raises a per-utterance flag whenever fp32 range may not
// have been enough, and flagged utterances are recomputed by the fp64 kernel (loss_launch, below).
//
// Why a second kernel.  ncu showed the fp64 kernel to be bound by its instruction count (240 M
// warp-instructions at config 2) with the lattice products sitting on the warp that owns the serial
// chain.  This one keeps the serial chains bare and moves everything else to block-parallel workers:
//
//   MF / MB   main sweeps (1 warp each): alpha over frames 0..T-1, beta over T-1..0, concurrently.
//             A lane owns NC consecutive lattice cells in registers (block floating point: fp32
//             values + one integer exponent per lane).  Every F_K frames (a "block" boundary) the
//             strip is renormalised by an exact power of two and written to an L2-resident
//             checkpoint.  Nothing else is on the chain: no products, no stores per frame.
//   LF / LB   loaders (1 warp each): 1-D TMA bulk copies (cp.async.bulk + mbarrier) of the next
//             block's log-prob rows, converted to probabilities for their main.
//   W0..W3    workers: pull blocks, middle first (a block is ready once MF has passed its start and
//             MB its end).  Per block: re-run beta from its checkpoint backwards over the block's
//             frames keeping the pre-emission sums in shared memory, re-run alpha forwards from
//             its checkpoint, multiply (registers), reduce the posteriors per label class (staging is
//             sorted by class: one warp prefix scan, class sum = difference) and write
//             scale * (p - posterior).  Blocks are independent, so workers never wait on each other.
//
// alpha and beta never reach HBM; traffic = log-probs + gradient + the per-block checkpoints.
//
// Range and the flag.  Linear-domain fp32 cannot represent every lattice (cells of one lane may
// legitimately differ by more than 2^150 when confident posteriors disagree with the target).
// Instead of proving range a priori the kernel checks the result: any mass lost to underflow in
// either sweep makes sum_v posterior_t(v) differ between frames, so every frame's sum is compared
// with 1 (D3).  Together with D1 (a log-prob beyond +-80 or NaN was seen: exp leaves fp32 range), D2 (a strip went
// non-finite) and D4 (Z not a positive finite number: infeasible or lost) this decides the flag.
#include <cmath>
#include <cstdlib>

#include "ssd_common.cuh"

// per-utterance set-up loops (executed once per utterance): rolled when SSD_F_SETUP_UNROLL1 is set, to keep them out of
// the instruction cache's way (experiment knob)
#ifndef SSD_F_SETUP_UNROLL1
#define SSD_F_SETUP_UNROLL1 0
#endif
#if SSD_F_SETUP_UNROLL1 && !defined(SSD_EMU)
#define SSD_F_SETUP_ROLL _Pragma("unroll 1")
#else
#define SSD_F_SETUP_ROLL
#endif

namespace ssd {

constexpr int F_K = 8;          // frames per block = renormalisation and checkpoint interval
constexpr int F_KS = 9;         // class-row stride of a probability block in floats: odd, so that a warp's gather of 32
                                // different classes at one frame touches 32 different banks (stride 8 folds them onto 4)
// Worker shape.  Default: 4 workers, whole blocks (8 warps, 128 registers).  -DSSD_F_W=6 -DSSD_F_H=4 selects the
// other point tried: 6 workers whose row buffers hold half a block (the block is done as two passes of 4 frames,
// the upper half's beta swept once more without stores); its 10 warps cap registers at 96 and the spills made
// it slower (562 us vs 393 us at config 2), see DESIGN.md.
#ifndef SSD_F_W
#define SSD_F_W 4
#endif
#ifndef SSD_F_H
#define SSD_F_H 8
#endif
constexpr int F_W = SSD_F_W;      // block workers
constexpr int F_THREADS = 32 * (4 + F_W);  // MF MB LF LB + workers
constexpr int F_H = SSD_F_H;      // frames per worker pass (rows a worker buffers)
static_assert((F_W == 4 || F_W == 6) && (F_H == 4 || F_H == 8), "supported worker shapes");
#ifndef SSD_F_NP
#define SSD_F_NP 3
#endif
#ifndef SSD_F_NRAW
#define SSD_F_NRAW 3
#endif
constexpr int F_NP = SSD_F_NP;      // probability-row ring depth per main (blocks)
constexpr int F_NRAW = SSD_F_NRAW;  // raw-row ring depth per main
#ifndef SSD_F_EARLYA
#define SSD_F_EARLYA 0  // 1: a worker loads the alpha checkpoint together with the beta one (held in registers through the beta pass)
#endif
#ifndef SSD_F_CU
#define SSD_F_CU 1  // unroll of the log-prob -> probability conversion loop (2V/32 trips)
#endif
#ifndef SSD_F_CP
#define SSD_F_CP 2  // class passes summed per trip of the range-sum loop (1 or 2)
#endif
#ifndef SSD_F_KU
#define SSD_F_KU 2  // unroll of the class range-sum loop
#endif
#define SSD_PRAGMA_(x) _Pragma(#x)
#define SSD_UNROLL(n) SSD_PRAGMA_(unroll n)
#ifndef SSD_F_FR
#define SSD_F_FR 4
#endif
#ifndef SSD_X
#define SSD_X 0  // timing experiments only (profiles/build_variant.sh): 1 no checkpoint stores, 2 no renormalisation, 4 no sweep steps, 8 no class sums, 16 no gradient stores, 32 no mass check
#endif
#ifndef SSD_F_PFD
#define SSD_F_PFD 0  // L2 prefetch distance of the loaders, in blocks (0 = off: no gain measured, warm or cold)
#endif
#ifndef SSD_F_L2HINT
#define SSD_F_L2HINT 1  // L2 eviction priorities on the bulk copies of the log-prob rows (keep for re-reads, drop after the last)
#endif
#ifndef SSD_F_LDSLEEP
#define SSD_F_LDSLEEP 2000u  // ns between a loader's polls of its ring's empty barrier; 0 = hardware-suspended try_wait
#endif
#ifndef SSD_F_PAIR
#define SSD_F_PAIR 1  // frame loops: 1 = two steps per trip with alternating probability registers, 0 = one step per trip
#endif
constexpr int F_FR = SSD_F_FR;  // frames interleaved in the gradient stage
constexpr int F_TOP = 40;  // renormalised strips have their largest cell in [2^40, 2^41)
constexpr int F_SPAD = 256;  // label staging capacity (S <= 255 on this path)
// Strip widths compiled in, one kernel each.  An utterance runs at the smallest one that holds its 2S+1 cells: sweeps,
// block rows and products all scale with it (config 2: labels U[50,200] need 4..13 cells per lane, 9 on average; one
// launch-wide width made every utterance pay for the longest).
constexpr int F_NCLS = 7;
__host__ __device__ inline int fast_class_of(int cells_per_lane) {
  return cells_per_lane <= 4 ? 0 : cells_per_lane <= 6 ? 1 : cells_per_lane <= 8 ? 2 : cells_per_lane <= 10 ? 3
       : cells_per_lane <= 12 ? 4 : cells_per_lane <= 14 ? 5 : 6;
}
__host__ __device__ inline int fast_width_of(int cls) { return 4 + 2 * cls; }
constexpr unsigned F_YIELD_NS = 3000u;  // head start given to CTAs on idle SMs at the first draw of a kernel
constexpr int F_RANK_MAX = 1024;  // batches up to this size are ranked by cost within their class (one CTA, O(B^2 / 1024))
#ifndef SSD_F_REBASE
#define SSD_F_REBASE 32
#endif
constexpr int F_REBASE = SSD_F_REBASE;  // a live neighbour this many binades above me: take its exponent (my cells shrink, its inflow stays finite)
// D3's tolerance on a frame's posterior mass.  What the result is held to is 1e-5 of max |gradient|, and a mass error of
// m shows up as about m in the posteriors: 1.5e-5 (round 1) let a row through that was 2.5e-5 off (round-2 fuzz: one
// label on 176 confident frames).  Ordinary rows sit at 1-3e-6: 6e-6 flags nothing at config 2, flat or trained.
#ifndef SSD_F_MASS_TOL
#define SSD_F_MASS_TOL 6e-6f
#endif
constexpr float F_MASS_TOL = SSD_F_MASS_TOL;
constexpr float F_LP_ABS = 80.f;  // log-probs beyond +-80 (or NaN) send the utterance to the exact path

// ------------------------------------------------------------------ scalar helpers
// 2^d for d in [-126, 127]; zero below, 2^127 above
__device__ __forceinline__ float pow2c(int d) {
  d = d > 127 ? 127 : d;
  return d < -126 ? 0.f : __int_as_float((d + 127) << 23);
}
__device__ __forceinline__ int expo_f(float v) { return ((__float_as_int(v) >> 23) & 0xff) - 127; }
__device__ __forceinline__ float ex2_approx(float x) {
#ifdef SSD_EMU
  return exp2f(x);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
// exp(lp) = ex2(lp * log2 e), with the product's rounding error put back: t = lp * L2E carries |t| * 2^-24 of error in
// the EXPONENT, i.e. a relative error of |t| * 4e-8 in p (1e-6 at lp = -17, the level of an unlikely label under
// confident posteriors).  Over hundreds of frames of a tiny lattice that perturbation showed up as 1.4e-5 in the
// gradient (round-2 fuzz: one label on 600 confident frames), above the 1e-5 the result is held to.  The residual of
// the product (one FMA) plus the low half of log2 e goes in as a first-order correction: what is left is ex2.approx's
// own 2^-22.  Inputs beyond +-F_LP_ABS are flagged by the caller (D1) and never reach a result.
__device__ __forceinline__ float prob_f(float lp) {
  const float L2E_HI = 1.44269502162933349609375f, L2E_LO = 1.925963033500011e-08f;
  const float t = lp * L2E_HI;
  const float d = fmaf(lp, L2E_LO, fmaf(lp, L2E_HI, -t));  // what t misses of lp * log2 e
  const float q = ex2_approx(t);
  return fmaf(q * 0.6931471805599453f, d, q);  // 2^(t+d) = q * (1 + d ln 2) to first order (|d| < 2^-20)
}
__device__ __forceinline__ float ld_cg(const float* p) {
#ifdef SSD_EMU
  return *p;
#else
  return __ldcg(p);
#endif
}
__device__ __forceinline__ int ld_cg(const int* p) {
#ifdef SSD_EMU
  return *p;
#else
  return __ldcg(p);
#endif
}

// Sum over the warp of values in [0, 1.5]: 2^-26 fixed point through the integer reduction unit (one REDUX
// instead of five dependent shuffle rounds; 32 * 1.5 * 2^26 < 2^32).  A lane holding anything else (NaN, inf,
// negative, larger) counts as 1.5, which no valid frame can absorb: D3 fires.
__device__ __forceinline__ float warp_sum_fx(float v) {
  const bool valid = v >= 0.f && v <= 1.5f;
  const unsigned q = (unsigned)__float2int_rn((valid ? v : 1.5f) * 67108864.f);
  return (float)__reduce_add_sync(FULL, q) * (1.f / 67108864.f);
}

struct FastParams {
  const float* lp;
  const float* lse;    // non-null: `lp` holds LOGITS and lse[b * lse_stride + t] their per-frame log-sum-exp (row f-2)
  int64_t lse_stride;  // multiple of F_K: a block's F_K values are one aligned 32-byte bulk copy
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
  unsigned* counter;   // [1] number of flagged utterances, [2] checkpoint regions handed out, [16 + 4c ..] class c: draws, head, tail
  unsigned* sm_slot;   // [#SM] arrival counter of working CTAs per SM (role rotation, dealing); [512 + 2 smid + s] region s of SM busy
  // written by loss_fast_prepass_kernel: utterances grouped by strip-width class (widest first), by decreasing cost within
  const int* cls_list;   // [B]
  const int* cls_start;  // [F_NCLS] first position of class c in cls_list
  const int* cls_count;  // [F_NCLS]
  int ranked;            // cls_list is cost-ordered within a class: the second CTA of an SM draws from the cheap end
  int region_by_sm;      // 1: more utterances than checkpoint regions, a CTA takes one of its SM's two regions and returns it
  int pdl_wait;          // first class kernel after the prepass: wait for it (programmatic dependent launch)
  int direct;            // 1: no prepass -- one kernel takes all B utterances in the caller's order (B <= checkpoint regions)
  int* flags;          // [B] non-zero = recompute on the exact path (1 not covered, 2 D1, 4 D2, 8 D3, 16 D4, 32 D5)
  unsigned char* ckpt; // per CTA: [2 dirs][nb_max] records
  size_t ckpt_cta_bytes, ckpt_rec_bytes;
  int nb_max;
  int RS;  // worker row stride in floats
  int use_tma;
  int rotate;  // 1: the second CTA of an SM mirrors its role map (tuning knob SSD_LOSS_NOROT)
  int late_worker;  // how many more workers (0..2) the sweeps' ring space hosts once both sweeps are done (SSD_LOSS_NOLATE)
  long long* tlog;  // SSD_LOSS_DBG & 8: [F_THREADS/32 warps][F_TLOG_N events][F_TLOG_S stamps] clock64 of CTA 0's first utterance
};
constexpr int F_TLOG_N = 256, F_TLOG_S = 6;

// compiled in only with -DSSD_LOSS_TIMELINE (profiles/timeline_fast.py builds that variant)
__device__ __forceinline__ void fstamp(const FastParams& P, unsigned uiter, int idx, int slot, long long tag = -1) {
#if !defined(SSD_EMU) && defined(SSD_LOSS_TIMELINE)
  if (P.tlog && blockIdx.x == 0 && uiter == 0 && idx < F_TLOG_N && (threadIdx.x & 31) == 0)
    P.tlog[(((int)threadIdx.x >> 5) * F_TLOG_N + idx) * F_TLOG_S + slot] = tag >= 0 ? tag : clock64();
#endif
}

struct FastSmem {  // byte offsets into dynamic shared memory, identical on host and device
  size_t p_slot, raw_slot, main_bytes, wrow_bytes, worker_bytes;
  size_t main_off, worker_off, lab_off, pos_off, cls_off, tab_off, bar_off, blk_off, scal_off, total;
};

// mbarrier indices: per main F_NP full + F_NP empty + F_NRAW raw; then z_ready; then 2 raw barriers per worker
constexpr int FB_PFULL = 0, FB_PEMPTY = F_NP, FB_RAW = 2 * F_NP, FB_PER_MAIN = 2 * F_NP + F_NRAW;
constexpr int FB_Z = 2 * FB_PER_MAIN, FB_WRAW = FB_Z + 1, FB_TOTAL = FB_WRAW + 2 * (F_W + 2);  // + the late workers' pairs

// Probability blocks are stored class-major: p[v * F_KS + r] for the block's frame r (row V is the all-zero dummy
// class of out-of-lattice cells).  A lane's label column is then ONE pointer for the whole block and a frame is
// an immediate offset from it.
__host__ __device__ inline FastSmem fast_smem_layout(int V, int RS, int S_tab, int nb_max) {
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  FastSmem m;
  m.p_slot = al((size_t)(V + 1) * F_KS * 4 + 16);
  m.raw_slot = al((size_t)F_K * V * 4 + F_K * 4);  // + the block's per-frame log-sum-exp values (logits input)
  m.main_bytes = F_NP * m.p_slot + F_NRAW * m.raw_slot;
  m.wrow_bytes = al((size_t)F_H * RS * 4);
  m.worker_bytes = m.wrow_bytes + m.p_slot + 2 * m.raw_slot;
  size_t o = 0;
  m.main_off = o; o += 2 * m.main_bytes;
  m.worker_off = o; o += F_W * m.worker_bytes;
  m.lab_off = o; o += al((size_t)S_tab * 4);
  m.pos_off = o; o += al((size_t)S_tab * 4);
  m.cls_off = o; o += al((size_t)(V + 2) * 4);
  m.tab_off = o; o += 4 * 16 * 32 * 4;  // per-lane label tables, see build_lane_tables
  m.bar_off = o; o += al((size_t)FB_TOTAL * 8);
  m.blk_off = o; o += al((size_t)nb_max * 8);
  m.scal_off = o; o += 128;
  m.total = o;
  return m;
}

struct FastScal {
  int utt, b, T, S, skip;
  float scale;
  int next_blk;    // worker block queue
  int uflag;       // this utterance needs the exact path (reason bits)
  float inv_zm;    // 1 / mantissa of Z
  int ez;          // exponent of Z
  int rot;         // role rotation of this CTA (0/1)
  int second;      // this CTA was the second working one to arrive on its SM
  int region;      // checkpoint region of this CTA
};

// ------------------------------------------------------------------ strips
// Strip of NC cells per lane (cell = lane*NC + i, even i = blank), true value = x[i] * 2^e.
template <int NC>
struct FStrip {
  float x[NC];
  // 2^(neighbour e - my e) as a product of two powers of two: brings the neighbour's boundary cells into my basis.
  // One float cannot hold it: adjacent lanes of a short, steep lattice differ by more than 2^126 while the
  // boundary cell times the factor is still a normal number (and 3 % of its neighbour cell, not noise).
  float fin, fin2;
  int e;
};

// per-lane label tables of one sweep direction: offset of every label cell's class in a probability
// block, 0/1 skip multipliers, optionally the class-sorted staging slot
template <int NC>
struct FLabels {
  int off[NC / 2];
  float skf[NC / 2];
};

// The tables are the same for every warp of the CTA: built once per utterance into shared memory as
// [4][NC/2][32] (class offset | alpha skip | beta skip | class-sorted staging slot), one entry per thread,
// and read by each role when (and only while) it needs them.
template <int NC>
__device__ __forceinline__ void build_lane_tables(int S, int V, const int* __restrict__ lab, const int* __restrict__ pos, int* tab) {
  const int t = (int)threadIdx.x;
  if (t >= (NC / 2) * 32) return;
  const int k = t >> 5, lane = t & 31;
  const int L = 2 * S + 1;
  const int c = lane * NC + 2 * k + 1;  // odd cell
  int col = V;                          // dummy class (p = 0) outside the lattice
  float ska = 0.f, skb = 0.f;
  int sp = -1;
  if (c < L) {
    const int j = (c - 1) >> 1;
    col = lab[j];
    sp = pos[j];
    if (c >= 3 && lab[j - 1] != col) ska = 1.f;
    if (c + 2 < L && lab[j + 1] != col) skb = 1.f;
  }
  constexpr int N = (NC / 2) * 32;
  tab[t] = col * F_KS * 4;  // byte offset of the class row in a probability block
  tab[N + t] = __float_as_int(ska);
  tab[2 * N + t] = __float_as_int(skb);
  tab[3 * N + t] = sp;
}
template <int NC, int DIR>
__device__ __forceinline__ void load_labels(FLabels<NC>& lt, const int* __restrict__ tab) {
  const int lane = lane_id();
  constexpr int N = (NC / 2) * 32;
#pragma unroll
  for (int k = 0; k < NC / 2; ++k) {
    lt.off[k] = tab[k * 32 + lane];
    lt.skf[k] = __int_as_float(tab[(DIR ? 2 : 1) * N + k * 32 + lane]);
  }
}

template <int NC>
struct FProbs {  // one frame
  float lab[NC / 2];
  float blank;
};
// Frame iteration over a class-major block.  Each label cell keeps ONE byte pointer into the block (its class
// row); a frame is that pointer plus an immediate, so a step costs no address arithmetic.  Two frames per
// trip with alternating register sets: the next frame's probabilities are always in flight during a step and
// never copied.  body(r, probs) is called for the block's frames in sweep order.
#ifdef SSD_EMU
typedef const char* saddr_t;
__device__ __forceinline__ saddr_t saddr_of(const void* p) { return reinterpret_cast<const char*>(p); }
template <int IMM>
__device__ __forceinline__ float lds_pinned(saddr_t a) {
  SSD_ASSERT((const unsigned char*)(a + IMM) >= emu::dyn_smem() && (const unsigned char*)(a + IMM) + 4 <= emu::dyn_smem() + emu::dyn_smem_size());
  return *reinterpret_cast<const float*>(a + IMM);
}
__device__ __forceinline__ saddr_t sadd(saddr_t a, int d) { return a + d; }
#else
typedef uint32_t saddr_t;
__device__ __forceinline__ saddr_t saddr_of(const void* p) { return smem_u32(p); }
// volatile: the compiler may not sink the load into the branch that consumes it, which is what keeps the
// next frame's probabilities in flight one whole step ahead
template <int IMM>
__device__ __forceinline__ float lds_pinned(saddr_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(IMM));
  return v;
}
__device__ __forceinline__ saddr_t sadd(saddr_t a, int d) { return a + (uint32_t)d; }
#endif
template <int NC>
struct PPtr {
  saddr_t lab[NC / 2];
  saddr_t blank;
};
template <int NC, int IMM>
__device__ __forceinline__ void load_probs(FProbs<NC>& p, const PPtr<NC>& pp) {
  p.blank = lds_pinned<IMM>(pp.blank);
#pragma unroll
  for (int k = 0; k < NC / 2; ++k) p.lab[k] = lds_pinned<IMM>(pp.lab[k]);
}
template <int NC, int DIR, class F>
__device__ __forceinline__ void for_frames(const float* pblk, int f0, int n, const FLabels<NC>& lt, int blank_off, F&& body) {
  constexpr int D = DIR ? -4 : 4;
  const saddr_t base = sadd(saddr_of(pblk), (DIR ? f0 + n - 1 : f0) * 4);
  PPtr<NC> pp;
  pp.blank = sadd(base, blank_off);
#pragma unroll
  for (int k = 0; k < NC / 2; ++k) pp.lab[k] = sadd(base, lt.off[k]);
#if SSD_F_PAIR
  FProbs<NC> A, B;
  load_probs<NC, 0>(A, pp);
#pragma unroll 1
  for (int rr = 0; rr < n; rr += 2) {
    const int r = DIR ? f0 + n - 1 - rr : f0 + rr;
    const bool two = rr + 1 < n;
    const bool more = rr + 2 < n;
    // each step gets the other register set's loads as its mid-step hook
    body(r, A, [&]() {
      if (two) load_probs<NC, D>(B, pp);
    });
    if (two) {
      body(DIR ? r - 1 : r + 1, B, [&]() {
        if (more) {
          pp.blank = sadd(pp.blank, 2 * D);
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) pp.lab[k] = sadd(pp.lab[k], 2 * D);
          load_probs<NC, 0>(A, pp);
        }
      });
    }
  }
#else
  // one step per trip: half the code of the two-set version at the price of NC/2+1 register moves per step
  FProbs<NC> A;
  load_probs<NC, 0>(A, pp);
#pragma unroll 1
  for (int rr = 0; rr < n; ++rr) {
    const int r = DIR ? f0 + n - 1 - rr : f0 + rr;
    FProbs<NC> B;
    body(r, A, [&]() {
      if (rr + 1 < n) {
        pp.blank = sadd(pp.blank, D);
#pragma unroll
        for (int k = 0; k < NC / 2; ++k) pp.lab[k] = sadd(pp.lab[k], D);
        load_probs<NC, 0>(B, pp);
      }
    });
    A = B;
  }
#endif
}

// One frame of the recursion.  DIR 0 (alpha): cell s pulls from s-1 and, label cells with a different
// label two back, s-2.  DIR 1 (beta): the mirror image, from s+1 and s+2.  SUMS: also return the
// pre-emission sums (what the posterior needs from the beta side).
//
// The neighbour lane's boundary cells arrive by shuffle.  A shuffle costs an in-order warp ~25 cycles twice
// (result latency, and the source register may not be overwritten before it is read), so they are
// software-pipelined: a step first computes the cells the NEXT step's shuffle reads, issues that shuffle,
// and only touches the values shuffled one step earlier (`in`) at its very end.
struct Inflow {
  float a, b;  // neighbour's boundary cells in ITS basis (b: beta only)
};
template <int NC, int DIR>
__device__ __forceinline__ void inflow_prime(const FStrip<NC>& s, Inflow& in) {
  if (DIR == 0) {
    in.a = __shfl_up_sync(FULL, s.x[NC - 1], 1);
    in.b = 0.f;
  } else {
    in.a = __shfl_down_sync(FULL, s.x[0], 1);
    in.b = __shfl_down_sync(FULL, s.x[1], 1);
  }
}
struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};
// `mid` runs right after the shuffles are issued: the place for the next frame's shared-memory loads, which
// would otherwise sit in the memory pipe in front of the shuffles and delay them
template <int NC, int DIR, bool SUMS, class H = NoHook>
__device__ __forceinline__ void fstep(FStrip<NC>& s, Inflow& in, const FLabels<NC>& lt, const FProbs<NC>& p, float* sums, H&& mid = NoHook()) {
  float (&x)[NC] = s.x;
  Inflow nxt;
  auto cell = [&](int i, float a1, float a2) {
    float sum = x[i] + a1;
    if (i & 1) {
      sum = fmaf(lt.skf[i >> 1], a2, sum);
      if (SUMS) sums[i] = sum;
      x[i] = sum * p.lab[i >> 1];
    } else {
      if (SUMS) sums[i] = sum;
      x[i] = sum * p.blank;
    }
  };
  if (DIR == 0) {
    // descending: inputs of lower cells are still intact
    cell(NC - 1, x[NC - 2], x[NC - 3]);
    nxt.a = __shfl_up_sync(FULL, x[NC - 1], 1);
    nxt.b = 0.f;
    mid();
#pragma unroll
    for (int i = NC - 2; i >= 2; --i) cell(i, x[i - 1], x[i - 2]);
    const float n1 = in.a * s.fin * s.fin2;  // lane 0: fin = 0
    cell(1, x[0], n1);
    cell(0, n1, 0.f);
  } else {
    // ascending
    cell(0, x[1], 0.f);
    cell(1, x[2], x[3]);
    nxt.a = __shfl_down_sync(FULL, x[0], 1);
    nxt.b = __shfl_down_sync(FULL, x[1], 1);
    mid();
#pragma unroll
    for (int i = 2; i < NC - 1; ++i) cell(i, x[i + 1], x[i + 2 < NC ? i + 2 : NC - 1]);
    const float m0 = in.a * s.fin * s.fin2, m1 = in.b * s.fin * s.fin2;  // lane 31: fin = 0
    cell(NC - 1, m0, m1);
  }
  in = nxt;
}

// how much of a lane's posterior scale 2^qe the alpha side takes (the beta~ side takes the rest): half, but no more
// than either side can hold next to values of up to 2^(TOP+14)
__device__ __forceinline__ int scale_split(int qe) {
  int h = qe >> 1;
  const int cap = 126 - (F_TOP + 14 + F_REBASE);  // a strip's cells reach 2^(TOP+14) on their own, times the inflow factor
  h = h < qe - cap ? qe - cap : h;
  return h > cap ? cap : h;
}

// the boundary factor both the mains and the workers derive from the lane exponents
template <int NC, int DIR>
__device__ __forceinline__ void set_fin(FStrip<NC>& s, int e_me) {
  const int lane = lane_id();
  const int e_nb = DIR == 0 ? __shfl_up_sync(FULL, e_me, 1) : __shfl_down_sync(FULL, e_me, 1);
  const bool edge = DIR == 0 ? lane == 0 : lane == 31;
  int d = e_nb - e_me;
  d = d < -252 ? -252 : (d > 252 ? 252 : d);
  const int d1 = d >> 1;
  s.fin = edge ? 0.f : pow2c(d1);
  s.fin2 = pow2c(d - d1);
}

// Block boundary: bring my largest cell to [2^TOP, 2^(TOP+1)), agree on exponents with the
// neighbour the mass comes from.  Returns false if the strip is not finite (D2).
template <int NC, int DIR>
__device__ __forceinline__ bool renorm_exchange(FStrip<NC>& s) {
  const int lane = lane_id();
  float m = s.x[0];
#pragma unroll
  for (int i = 1; i < NC; ++i) m = fmaxf(m, s.x[i]);
  const bool ok = m < 3.0e38f;              // false for inf; NaN cells surface through D3/D4
  const bool live = m >= 1.17549435e-38f;   // below the normal range: treat as empty
  int k = expo_f(m) - F_TOP;
  k = k < -120 ? -120 : k;
  k = live ? k : 0;
  const float f = live ? pow2c(-k) : 0.f;
#pragma unroll
  for (int i = 0; i < NC; ++i) s.x[i] *= f;
  s.e += k;
  // Strips DOWNSTREAM of the strip with the largest basis (the mass front, or the bulk) whose own basis is more than
  // F_REBASE binades below it take that basis: whatever reaches them during the coming block comes from there, and
  // an inflow factor of at most 2^F_REBASE is what the products downstream (Z, posteriors) are dimensioned for.
  // Their own cells shrink by an exact power of two; what underflows was 2^-166 of the front.  One REDUX and one
  // ballot, no loop: with confident posteriors the front is a cliff of ~2^150 per strip, several strips wide.
  const unsigned lv = __ballot_sync(FULL, live);
  {
    const int G = __reduce_max_sync(FULL, live ? s.e : -(1 << 30));
    const unsigned top = __ballot_sync(FULL, live && s.e == G);
    const bool down = DIR == 0 ? (top & ((1u << lane) - 1u)) != 0u : (lane < 31 && (top >> (lane + 1)) != 0u);
    const int gap = G - s.e;
    if (live && down && gap > F_REBASE) {
      const float g = pow2c(-gap);
#pragma unroll
      for (int i = 0; i < NC; ++i) s.x[i] *= g;
      s.e = G;
    }
  }
  // an empty strip takes the exponent of the nearest live strip on the side its mass will come from
  int src;
  bool fed;  // some strip on the side my mass comes from is live: my neighbour's exponent is that of real values
  if (DIR == 0) {
    const unsigned below = lv & ((1u << lane) - 1u);
    src = below ? 31 - __clz((int)below) : lane;
    fed = below != 0u;
  } else {
    const unsigned above = lane == 31 ? 0u : (lv >> (lane + 1)) << (lane + 1);
    src = above ? __ffs((int)above) - 1 : lane;
    fed = above != 0u;
  }
  const int e_near = __shfl_sync(FULL, s.e, src);
  if (!live) s.e = e_near;
  // What the rule above leaves (profiles with a second, lower hump): a neighbour on the feeding side whose basis
  // dwarfs mine -- rebase on it, so that whatever it hands over during the coming block stays finite in my basis.
  // The neighbour itself may be empty right now: it then carries the exponent of the live strip beyond it and is a
  // conduit (an empty strip fills up within the block; its outflow overflowed a strip that had kept a basis 2^199
  // lower: peaked posteriors on their own transcript, i.e. every trained model, were flagged for that).  A
  // neighbour with nothing live behind it is stale and never rebased on.  A strip that rebases widens its own
  // downstream neighbour's gap, so the test repeats until no strip moves: one vote and no work almost always.
  int e_nb = 0;
#pragma unroll 1
  for (int round = 0; round < 32; ++round) {
    e_nb = DIR == 0 ? __shfl_up_sync(FULL, s.e, 1) : __shfl_down_sync(FULL, s.e, 1);
    const int d = e_nb - s.e;
    const bool move = fed && d > F_REBASE;
    if (!__any_sync(FULL, move)) break;
    if (move) {
      const float g = pow2c(-d);
#pragma unroll
      for (int i = 0; i < NC; ++i) s.x[i] *= g;
      s.e = e_nb;
    }
  }
  set_fin<NC, DIR>(s, s.e);
  return ok;
}

template <int NC, int DIR>
__device__ __forceinline__ void sweep_init_f(int L, FStrip<NC>& s) {
  const int lane = lane_id();
  const int c0 = DIR ? L - 1 : 0;  // the virtual frame before the first one holds all mass in one cell
#pragma unroll
  for (int i = 0; i < NC; ++i) s.x[i] = (lane * NC + i == c0) ? 1.f : 0.f;
  s.e = 0;
  s.fin = (DIR == 0 ? lane == 0 : lane == 31) ? 0.f : 1.f;
  s.fin2 = 1.f;
}

// checkpoint record: [NC][32] cells, [32] exponents; lanes beyond the lattice hold zeros and are skipped
template <int NC>
__device__ __forceinline__ void ckpt_store(unsigned char* rec, const FStrip<NC>& s, int L) {
  const int lane = lane_id();
  float* cx = reinterpret_cast<float*>(rec);
  if (lane * NC < L) {
#pragma unroll
    for (int i = 0; i < NC; ++i) cx[i * 32 + lane] = s.x[i];
  }
  reinterpret_cast<int*>(cx + NC * 32)[lane] = s.e;
}
template <int NC>
__device__ __forceinline__ void ckpt_load(const unsigned char* rec, FStrip<NC>& s, int L) {
  const int lane = lane_id();
  const float* cx = reinterpret_cast<const float*>(rec);
  const bool in = lane * NC < L;
#pragma unroll
  for (int i = 0; i < NC; ++i) s.x[i] = in ? ld_cg(cx + i * 32 + lane) : 0.f;
  s.e = ld_cg(reinterpret_cast<const int*>(cx + NC * 32) + lane);
}

struct FPlan {  // per utterance
  int b, Tb, S, L, NB;
};

// i-th block in (approximate) order of readiness.  Block q is ready at step max(q*K, Tb - (q+1)*K): two
// arithmetic sequences with the same stride that meet at qc, so walking outwards from qc alternately
// visits them in order; whichever side runs out first leaves the rest to the other.
__device__ __forceinline__ int block_order(int i, int NB, int Tb) {
  int qc = (Tb - F_K + 2 * F_K - 1) / (2 * F_K);  // first block whose readiness is set by the forward sweep
  qc = qc < 0 ? 0 : (qc > NB - 1 ? NB - 1 : qc);
  const int nhi = NB - qc, nlo = qc;
  const int m = nhi < nlo ? nhi : nlo;
  if (i < 2 * m) return (i & 1) ? qc - 1 - (i >> 1) : qc + (i >> 1);
  return nhi > nlo ? qc + (i - m) : qc - 1 - (i - m);
}

// Convert one block of raw log-prob rows into a class-major probability block.  `raw` is the block in
// shared memory ([n][V], TMA-loaded) or, if null, `graw` the same rows in global memory (row stride st).
// One (frame quad, class) item per lane and pass: four independent load -> exp chains in flight, small code.
// Returns non-zero if a value outside +-F_LP_ABS (or NaN) was seen (D1).
template <bool SMEM, bool LOGITS>
__device__ __forceinline__ int convert_block_t(const float* __restrict__ raw, const float* __restrict__ graw, int64_t st, int n, int V,
                                               float* __restrict__ pblk, const float* __restrict__ lse) {
  const int lane = lane_id();
  float amax = 0.f;  // largest |log-prob| seen (a NaN passes here and poisons Z or a frame's mass instead)
  SSD_UNROLL(SSD_F_CU)
  for (int base = 0; base < 2 * V; base += 32) {
    int idx = base + lane;  // item = (frame quad, class): consecutive lanes, consecutive classes
    const bool live = idx < 2 * V;
    idx = live ? idx : 2 * V - 1;
    const int hq = idx >= V ? 1 : 0;
    const int v = idx - hq * V;
    float lp[4], pr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int r = hq * 4 + j;
      r = r < n ? r : n - 1;  // rows beyond a partial block: re-read the last one, masked below
      lp[j] = SMEM ? raw[r * V + v] : ld_stream(graw + (int64_t)r * st + v);
      // logits: log-prob = x - lse_t.  The F_K values sit behind the rows in the landing buffer (SMEM) or are read
      // from the global array
      if (LOGITS) lp[j] -= SMEM ? raw[F_K * V + r] : ld_cg(lse + r);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      amax = fmaxf(amax, fabsf(lp[j]));
      const float pv = prob_f(lp[j]);
      pr[j] = hq * 4 + j < n ? pv : 0.f;
    }
    if (live) {
      float* dst = pblk + v * F_KS + hq * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = pr[j];
    }
  }
  return amax > F_LP_ABS ? 1 : 0;
}
// one loop per source: the unused path's address arithmetic would otherwise ride along predicated off.
// NOT inlined: loaders and workers then run one shared copy, which is what the instruction cache wants
// (the kernel is instruction-fetch sensitive: four different roles issue from every sub-partition).
// `lse` non-null selects the logits variants (global pointer to the block's first frame; with `raw` set the
// values are taken from the landing buffer instead, where a second bulk copy has put them).
__device__ __noinline__ int convert_block(const float* __restrict__ raw, const float* __restrict__ graw, int64_t st, int n, int V,
                                             float* __restrict__ pblk, const float* __restrict__ lse) {
  if (lse) return raw ? convert_block_t<true, true>(raw, graw, st, n, V, pblk, lse) : convert_block_t<false, true>(raw, graw, st, n, V, pblk, lse);
  return raw ? convert_block_t<true, false>(raw, graw, st, n, V, pblk, lse) : convert_block_t<false, false>(raw, graw, st, n, V, pblk, lse);
}

struct MainView {  // one main's rings
  unsigned char* base;
  uint64_t* bar;
  size_t p_slot, raw_slot;
  __device__ __forceinline__ float* p(int slot) const { return reinterpret_cast<float*>(base + slot * p_slot); }
  __device__ __forceinline__ float* raw(int slot) const { return reinterpret_cast<float*>(base + F_NP * p_slot + slot * raw_slot); }
};

// ------------------------------------------------------------------ L: loader / converter of one main
template <int DIR>
__device__ void fast_loader(const FastParams& P, const MainView& mv, const FPlan& pl, unsigned qbase, FastScal* sc, unsigned uiter) {
  const int lane = lane_id();
  const int V = P.V;
  const float* base = P.lp + (int64_t)pl.b * P.sb;
  const float* lse_base = P.lse ? P.lse + (int64_t)pl.b * P.lse_stride : nullptr;
  auto blk_of = [&](int bi) { return DIR ? pl.NB - 1 - bi : bi; };
  auto issue_raw = [&](int bi) {  // lane 0
    const int q = blk_of(bi);
    const int n = min(F_K, pl.Tb - q * F_K);
    const int slot = (int)((qbase + (unsigned)bi) % F_NRAW);
    const uint32_t bytes = (uint32_t)(n * V * 4);
    mbar_arrive_expect_tx(&mv.bar[FB_RAW + slot], bytes + (lse_base ? F_K * 4u : 0u));
    // the sweeps' reads are the first two of three: ask L2 to keep the lines for the block workers
    if (SSD_F_L2HINT) bulk_g2s_hint(mv.raw(slot), base + (int64_t)q * F_K * P.st, bytes, &mv.bar[FB_RAW + slot], 1);
    else bulk_g2s(mv.raw(slot), base + (int64_t)q * F_K * P.st, bytes, &mv.bar[FB_RAW + slot]);
    if (lse_base) bulk_g2s(mv.raw(slot) + F_K * V, lse_base + q * F_K, F_K * 4u, &mv.bar[FB_RAW + slot]);
  };
  // The rows come from HBM the first time (~1.5 us per bulk copy, measured: with F_NRAW - 1 = 2 copies in flight the
  // ring delivered one block per ~1470 cycles and the sweeps starved).  A bulk L2 prefetch F_PFD blocks ahead costs
  // no shared memory and turns the ring's copies into L2 hits.
  auto prefetch = [&](int bi) {  // lane 0
    const int q = blk_of(bi);
    const int n = min(F_K, pl.Tb - q * F_K);
    bulk_prefetch_l2(base + (int64_t)q * F_K * P.st, (uint32_t)(n * V * 4));
  };
  if (P.use_tma && lane == 0) {
    for (int bi = 0; bi < pl.NB && bi < F_NRAW - 1; ++bi) issue_raw(bi);
    if (SSD_F_PFD)
      for (int bi = F_NRAW - 1; bi < pl.NB && bi < SSD_F_PFD; ++bi) prefetch(bi);
  }
  int lpbad = 0;
  for (int bi = 0; bi < pl.NB; ++bi) {
    __syncwarp();  // every lane is done with raw slot (bi-1) before lane 0 re-arms it
    if (P.use_tma && lane == 0) {
      if (bi + F_NRAW - 1 < pl.NB) issue_raw(bi + F_NRAW - 1);
      if (SSD_F_PFD && bi + SSD_F_PFD < pl.NB) prefetch(bi + SSD_F_PFD);
    }
    const unsigned Q = qbase + (unsigned)bi;
    const int slot = (int)(Q % F_NP);
    fstamp(P, uiter, bi, 0);
#if SSD_F_LDSLEEP
    mbar_wait_relaxed(&mv.bar[FB_PEMPTY + slot], ((Q / F_NP) & 1u) ^ 1u, SSD_F_LDSLEEP);  // I am blocks ahead: no hurry
#else
    mbar_wait(&mv.bar[FB_PEMPTY + slot], ((Q / F_NP) & 1u) ^ 1u);
#endif
    fstamp(P, uiter, bi, 1);
    const int q = blk_of(bi);
    const int n = min(F_K, pl.Tb - q * F_K);
    const float* raw = nullptr;
    if (P.use_tma) {
      const int rs = (int)(Q % F_NRAW);
      mbar_wait(&mv.bar[FB_RAW + rs], (Q / F_NRAW) & 1u);
      raw = mv.raw(rs);
    }
    fstamp(P, uiter, bi, 2);
    lpbad |= convert_block(raw, base + (int64_t)q * F_K * P.st, P.st, n, V, mv.p(slot), lse_base ? lse_base + q * F_K : nullptr);
    __syncwarp();
    if (lane == 0) mbar_arrive(&mv.bar[FB_PFULL + slot]);
    fstamp(P, uiter, bi, 3);
  }
  if (DIR == 0 && __any_sync(FULL, lpbad) && lane == 0) atomicOr(&sc->uflag, 2);  // D1
}

// ------------------------------------------------------------------ M: main sweep of one direction
template <int NC, int DIR>
__device__ void fast_main(const FastParams& P, const MainView& mv, const FPlan& pl, const int* __restrict__ tab, unsigned qbase,
                          unsigned char* ckpt_cta, uint64_t* blk_ready, FastScal* sc, unsigned uiter) {
  const int lane = lane_id();
  const int blank_off = P.blank * F_KS * 4;  // bytes
  FLabels<NC> lt;
  load_labels<NC, DIR>(lt, tab);
  FStrip<NC> s;
  sweep_init_f<NC, DIR>(pl.L, s);
  unsigned char* ck = ckpt_cta + (size_t)DIR * P.nb_max * P.ckpt_rec_bytes;
  // checkpoints: the lanes' stores are ordered before lane 0's (release) arrive by the __syncwarp
  {  // the start state is the checkpoint of my first block
    const int q0 = DIR ? pl.NB - 1 : 0;
    ckpt_store<NC>(ck + (size_t)q0 * P.ckpt_rec_bytes, s, pl.L);
    __syncwarp();
    if (lane == 0) mbar_arrive(&blk_ready[q0]);
  }
  Inflow in;
  inflow_prime<NC, DIR>(s, in);
  bool ok = true;
  // the workers sleep on named barrier 1 (no polling) until both sweeps are about a block away from meeting
  const int bi_wake = pl.NB / 2 > 1 ? pl.NB / 2 - 1 : 0;
  for (int bi = 0; bi < pl.NB; ++bi) {
    const int q = DIR ? pl.NB - 1 - bi : bi;
    const unsigned Q = qbase + (unsigned)bi;
    const int slot = (int)(Q % F_NP);
    if (bi == bi_wake) bar_arrive(1, 32 * (F_W + 2));
    fstamp(P, uiter, bi, 0);
    mbar_wait(&mv.bar[FB_PFULL + slot], (Q / F_NP) & 1u);
    fstamp(P, uiter, bi, 1);
    const float* pblk = mv.p(slot);
    const int n = min(F_K, pl.Tb - q * F_K);
    if (!(SSD_X & 4))
      for_frames<NC, DIR>(pblk, 0, n, lt, blank_off, [&](int, const FProbs<NC>& pc, auto&& mid) { fstep<NC, DIR, false>(s, in, lt, pc, nullptr, mid); });
    __syncwarp();
    if (lane == 0) mbar_arrive(&mv.bar[FB_PEMPTY + slot]);
    fstamp(P, uiter, bi, 2);
    if (!(SSD_X & 2)) {
      ok &= renorm_exchange<NC, DIR>(s);
      inflow_prime<NC, DIR>(s, in);  // the values shuffled before the renormalisation are in the old basis
    }
    const int qn = DIR ? q - 1 : q + 1;  // the block this state starts
    if (qn >= 0 && qn < pl.NB) {
      if (!(SSD_X & 1)) ckpt_store<NC>(ck + (size_t)qn * P.ckpt_rec_bytes, s, pl.L);
      __syncwarp();
      if (lane == 0) mbar_arrive(&blk_ready[qn]);
    }
    fstamp(P, uiter, bi, 3);
  }
  if (__any_sync(FULL, !ok) && lane == 0) atomicOr(&sc->uflag, 4);  // D2
}

// ------------------------------------------------------------------ W: block workers
struct WorkerView {
  float* rows;   // [F_H][RS]: beta~ rows of the half in flight, then (same storage) the staged posteriors
  float* pblk;   // class-major probabilities of the block in flight
  float* raw0;   // two TMA landing buffers (one block ahead), raw_stride floats apart
  int raw_stride;
  uint64_t* rawbar;  // [2]
};

template <int NC>
__device__ void fast_worker(const FastParams& P, const FPlan& pl, const int* __restrict__ tab,
                            const int* __restrict__ cls, const WorkerView& wv, const unsigned char* ckpt_cta, uint64_t* blk_ready,
                            uint64_t* z_ready, unsigned uiter, unsigned& nraw, FastScal* sc, bool late) {
  const int lane = lane_id();
  const int V = P.V, RS = P.RS;
  const int blank_off = P.blank * F_KS * 4;  // bytes
  const bool want_grad = P.grad != nullptr;
  float* rows = wv.rows;
  float* pblk = wv.pblk;
  const unsigned char* ck_a = ckpt_cta;
  const unsigned char* ck_b = ckpt_cta + (size_t)P.nb_max * P.ckpt_rec_bytes;
  const float* lp_base = P.lp + (int64_t)pl.b * P.sb;
  const float* lse_base = P.lse ? P.lse + (int64_t)pl.b * P.lse_stride : nullptr;
  float* g_base = want_grad ? P.grad + (int64_t)pl.b * P.sb : nullptr;
  const float scale = sc->scale;
  bool mass_ok = true;
  // class passes of the gradient stage: lane handles classes lane, lane+32, ..; per pass the longest class range
  const int ncls_it = (V + 31) >> 5;  // <= 4 (V <= 128)
  unsigned mlen = 0u;
  for (int ci = 0, v = lane; ci < ncls_it; ++ci, v += 32) {
    const int len = v < V ? cls[v + 1] - cls[v] : 0;
    mlen |= (unsigned)min(__reduce_max_sync(FULL, len), 255) << (8 * ci);
  }

  // without a gradient only Z is needed: the first block alone
  const int nblk = want_grad ? pl.NB : 1;
  auto grab = [&]() {
    int i = 0;
    if (lane == 0) i = atomicAdd(&sc->next_blk, 1);
    return __shfl_sync(FULL, i, 0);
  };
  auto issue_raw = [&](int i, unsigned use) {  // lane 0: the block's rows -> landing buffer (use & 1)
    const int q = block_order(i, pl.NB, pl.Tb);
    const int n = min(F_K, pl.Tb - q * F_K);
    const uint32_t bytes = (uint32_t)(n * V * 4);
    mbar_arrive_expect_tx(&wv.rawbar[use & 1u], bytes + (lse_base ? F_K * 4u : 0u));
    // the workers are the last readers of a block's rows
    if (SSD_F_L2HINT) bulk_g2s_hint(wv.raw0 + (use & 1u) * wv.raw_stride, lp_base + (int64_t)q * F_K * P.st, bytes, &wv.rawbar[use & 1u], 0);
    else bulk_g2s(wv.raw0 + (use & 1u) * wv.raw_stride, lp_base + (int64_t)q * F_K * P.st, bytes, &wv.rawbar[use & 1u]);
    if (lse_base) bulk_g2s(wv.raw0 + (use & 1u) * wv.raw_stride + F_K * V, lse_base + q * F_K, F_K * 4u, &wv.rawbar[use & 1u]);
  };
  if (!late) bar_sync(1, 32 * (F_W + 2));  // released by the two mains shortly before the first block can be ready
  int i = grab();
  if (i < nblk && P.use_tma && lane == 0) issue_raw(i, nraw);
  int nmine = 0;
  for (; i < nblk; ++nmine) {
    fstamp(P, uiter, nmine, 0);
    const int q = block_order(i, pl.NB, pl.Tb);
    const int t0 = q * F_K;
    const int n = min(F_K, pl.Tb - t0);
    // one block ahead: claim the next block and start its copy
    __syncwarp();  // the landing buffer about to be re-armed was read by all lanes one block ago
    const int i_next = grab();
    if (i_next < nblk && P.use_tma && lane == 0) issue_raw(i_next, nraw + 1);
    // ---- probabilities of the block's frames (own copy: workers run behind the mains' rings)
    const float* raw = nullptr;
    if (P.use_tma) {
      mbar_wait(&wv.rawbar[nraw & 1u], (nraw >> 1) & 1u);
      raw = wv.raw0 + (nraw & 1u) * wv.raw_stride;
    }
    ++nraw;
    convert_block(raw, lp_base + (int64_t)t0 * P.st, P.st, n, V, pblk, lse_base ? lse_base + t0 : nullptr);
    fstamp(P, uiter, nmine, 1);
    mbar_wait_relaxed(&blk_ready[q], uiter & 1u);  // both checkpoints of the block are published
    fstamp(P, uiter, nmine, 2);
    fstamp(P, uiter, nmine, 5, q);
    __syncwarp();
    // ---- the block in halves of F_H frames (lower half first: alpha runs upwards through both)
    FStrip<NC> sa;  // alpha state, carried from the lower half into the upper
    Inflow ina;
    int e_alpha = 0, e_beta = 0, g_beta = 0;
    bool stop = false;
    // every block but the first needs Z before its sweeps (they run pre-scaled); the first one computes it
    bool z_known = i != 0;
    if (z_known && want_grad) mbar_wait_relaxed(z_ready, uiter & 1u);
#pragma unroll 1
    for (int f0 = 0; f0 < n; f0 += F_H) {
      const int nh = min(F_H, n - f0);
      // ---- beta backwards: pre-emission sums of this half's frames -> rows.  Its start state is the block's
      // checkpoint swept (without stores) down through the frames above this half.
      {
        FStrip<NC> s;
        ckpt_load<NC>(ck_b + (size_t)q * P.ckpt_rec_bytes, s, pl.L);
#if SSD_F_EARLYA
        // the alpha checkpoint is not needed before the beta pass is over: its L2 latency hides behind that pass
        if (f0 == 0) ckpt_load<NC>(ck_a + (size_t)q * P.ckpt_rec_bytes, sa, pl.L);
#endif
        e_beta = s.e;
        if (f0 == 0) e_alpha = ld_cg(reinterpret_cast<const int*>(ck_a + (size_t)q * P.ckpt_rec_bytes + (size_t)NC * 32 * 4) + lane);
        if (z_known) {
          // posterior = alpha * beta~ * 2^(e_a + e_b - e_Z) / Zm.  The scale is SPLIT between the two recomputed sweeps
          // (alpha takes 2^h, beta~ takes 2^(qe - h) / Zm): either factor alone can leave the fp32 range when the two
          // sweeps have steep opposite slopes inside a lane, their product at a cell never does (it is a posterior).
          const int qe = e_alpha + e_beta - sc->ez;
          g_beta = qe - scale_split(qe);
          const float c = pow2c(g_beta) * sc->inv_zm;
#pragma unroll
          for (int k = 0; k < NC; ++k) s.x[k] *= c;
        } else {
          g_beta = 0;
        }
        set_fin<NC, 1>(s, s.e - g_beta);
        FLabels<NC> lb;
        load_labels<NC, 1>(lb, tab);
        Inflow in;
        inflow_prime<NC, 1>(s, in);
        if (F_H < F_K && f0 + nh < n)
          for_frames<NC, 1>(pblk, f0 + nh, n - f0 - nh, lb, blank_off,
                            [&](int, const FProbs<NC>& pc, auto&& mid) { fstep<NC, 1, false>(s, in, lb, pc, nullptr, mid); });
        for_frames<NC, 1>(pblk, f0, nh, lb, blank_off, [&](int r, const FProbs<NC>& pc, auto&& mid) {
          float sums[NC];
          fstep<NC, 1, true>(s, in, lb, pc, sums, mid);
          float2* dst = reinterpret_cast<float2*>(rows + (r - f0) * RS) + lane;
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) dst[k * 32] = make_float2(sums[2 * k], sums[2 * k + 1]);
        });
      }
      __syncwarp();
      fstamp(P, uiter, nmine, 3);
      // ---- alpha forwards
      FLabels<NC> la;
      load_labels<NC, 0>(la, tab);
      if (f0 == 0) {
#if !SSD_F_EARLYA
        ckpt_load<NC>(ck_a + (size_t)q * P.ckpt_rec_bytes, sa, pl.L);
#endif
        if (!z_known) {
          // Z = sum over the lattice of alpha_t * beta~_t at the block's first frame (unscaled pass)
          FStrip<NC> z = sa;
          set_fin<NC, 0>(z, z.e);
          Inflow zin;
          inflow_prime<NC, 0>(z, zin);
          FProbs<NC> pc;
          pc.blank = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(pblk) + blank_off);
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) pc.lab[k] = *reinterpret_cast<const float*>(reinterpret_cast<const char*>(pblk) + la.off[k]);
          fstep<NC, 0, false>(z, zin, la, pc, nullptr);
          const float2* src = reinterpret_cast<const float2*>(rows) + lane;
          // both factors may sit 2^F_REBASE above their strips' nominal top (inflow from a higher basis during the
          // block): each side is brought to a top of 2^F_TOP by an exact power of two before the products, and the
          // lane's exponent takes the two shifts back (once per utterance: the cost is nothing)
          float ma = 0.f, mb = 0.f;
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) {
            const float2 w = src[k * 32];
            ma = fmaxf(ma, fmaxf(z.x[2 * k], z.x[2 * k + 1]));
            mb = fmaxf(mb, fmaxf(w.x, w.y));
          }
          const int ka = ma >= 1.17549435e-38f && ma < 3.0e38f ? expo_f(ma) - F_TOP : 0;
          const int kb = mb >= 1.17549435e-38f && mb < 3.0e38f ? expo_f(mb) - F_TOP : 0;
          const float da = pow2c(-ka), db = pow2c(-kb);
          float sl = 0.f;
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) {
            const float2 w = src[k * 32];
            sl = fmaf(z.x[2 * k] * da, w.x * db, sl);
            sl = fmaf(z.x[2 * k + 1] * da, w.y * db, sl);
          }
          const bool okl = sl >= 1.17549435e-38f && sl < 3.0e38f;
          const int q_l = e_alpha + e_beta + ka + kb;
          const int key = okl ? q_l + expo_f(sl) : -(1 << 30);
          const int Qm = __reduce_max_sync(FULL, key);
          const bool bad_lane = !(sl >= 0.f && sl < 3.0e38f);  // NaN / inf somewhere
          const bool any_bad = __any_sync(FULL, bad_lane);
          float inv_zm = 0.f;
          int ez = 0;
          const bool feasible = Qm > -(1 << 30) && !any_bad;
          if (feasible) {
            // every lane's partial in units of 2^Qm: at most 2 each, negligible ones flush
            float tot = okl ? sl * pow2c(q_l - Qm) : 0.f;
#pragma unroll
            for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
            const int et = expo_f(tot);
            const float zm = tot * pow2c(-et);  // [1,2)
            inv_zm = 1.f / zm;
            ez = Qm + et;
            if (lane == 0) P.nll[pl.b] = (float)(-(log((double)zm) + (double)ez * 0.69314718055994530942));
            // D5: Z above 1/8, i.e. nll below ~2 (a handful of frames with confident posteriors).  Z carries fp32's
            // ~4e-7 relative error; -log amplifies it by 1/nll, and 1e-5 of the nll is what the result is held to.
            // Such an utterance is tiny: the exact path takes it.  (A training-length utterance has nll >> 2.)
            if (ez >= -3 && lane == 0) atomicOr(&sc->uflag, 32);
          } else if (lane == 0) {
            atomicOr(&sc->uflag, 16);  // D4: infeasible or lost -- the exact path decides which
          }
          if (lane == 0) {
            sc->inv_zm = inv_zm;
            sc->ez = ez;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(z_ready);
        }
        if (!want_grad) {
          stop = true;
          break;
        }
        const int qe = e_alpha + e_beta - sc->ez;
        const int h = scale_split(qe);
        if (!z_known) {
          // the first block's rows were stored unscaled: give them their share of the scale now (my own cells)
          z_known = true;
          g_beta = qe - h;
          const float c = pow2c(g_beta) * sc->inv_zm;
#pragma unroll 1
          for (int r = 0; r < nh; ++r) {
            float2* row = reinterpret_cast<float2*>(rows + r * RS) + lane;
#pragma unroll
            for (int k = 0; k < NC / 2; ++k) {
              float2 w = row[k * 32];
              w.x *= c;
              w.y *= c;
              row[k * 32] = w;
            }
          }
        }
        {
          const float c = pow2c(h);
#pragma unroll
          for (int k = 0; k < NC; ++k) sa.x[k] *= c;
        }
        set_fin<NC, 0>(sa, e_alpha - h);
        inflow_prime<NC, 0>(sa, ina);
      }
      {
        int spos[NC / 2];  // class-sorted staging slot of my label cells
#pragma unroll
        for (int k = 0; k < NC / 2; ++k) spos[k] = tab[3 * (NC / 2) * 32 + k * 32 + lane];
        for_frames<NC, 0>(pblk, f0, nh, la, blank_off, [&](int r, const FProbs<NC>& pc, auto&& mid) {
          float* gf = rows + (r - f0) * RS;
          const float2* src = reinterpret_cast<const float2*>(gf) + lane;
          float2 w[NC / 2];
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) w[k] = src[k * 32];
          fstep<NC, 0, false>(sa, ina, la, pc, nullptr, mid);
          __syncwarp();  // every lane has read the row before it becomes the staging area
          float gb = 0.f;
#pragma unroll
          for (int k = 0; k < NC / 2; ++k) {
            gb = fmaf(sa.x[2 * k], w[k].x, gb);
            if (spos[k] >= 0) gf[spos[k]] = sa.x[2 * k + 1] * w[k].y;
          }
          gf[F_SPAD + lane] = gb;
          if (lane == 0) gf[F_SPAD + 32] = 0.f;  // the always-zero slot out-of-range class reads are pointed at
        });
      }
      __syncwarp();
      fstamp(P, uiter, nmine, 4);
      // ---- class sums of the staged posteriors and gradient rows of this half's frames, all interleaved.
      // Staging is sorted by class, so a class is one contiguous range [cls[v], cls[v+1]): each lane adds up
      // the ranges of its classes directly (independent loads, no scan).  The two warp-wide sums per frame (blank
      // partials, total mass) go through the integer reduction unit in 2^-26 fixed point: one REDUX each
      // instead of five dependent shuffle rounds.
      constexpr int FR = F_FR;
#pragma unroll 1
      for (int r0 = 0; r0 < nh; r0 += FR) {
        const int nf = nh - r0 < FR ? nh - r0 : FR;
        const float* gfp[FR];
        float gbl[FR], gb[FR];
#pragma unroll
        for (int f = 0; f < FR; ++f) {
          gfp[f] = rows + (r0 + (f < nf ? f : 0)) * RS;
          gbl[f] = gfp[f][F_SPAD + lane];
        }
#pragma unroll
        for (int f = 0; f < FR; ++f) gb[f] = warp_sum_fx(gbl[f]);
        float tot[FR];
#pragma unroll
        for (int f = 0; f < FR; ++f) tot[f] = gbl[f];
        constexpr int CP = SSD_F_CP;  // class passes (32 classes each) summed together: more independent loads per trip
#pragma unroll 1
        for (int ci = 0; ci < ncls_it; ci += CP) {
          float acc[CP][FR];
          int ca[CP], cb[CP];
          int ml = 0;
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            const int v = lane + 32 * (ci + c);
            ca[c] = v < V ? cls[v] : 0;
            cb[c] = v < V ? cls[v + 1] : 0;
            ml = max(ml, ci + c < ncls_it ? (int)((mlen >> (8 * (ci + c))) & 0xff) : 0);  // longest range among these classes
#pragma unroll
            for (int f = 0; f < FR; ++f) acc[c][f] = 0.f;
          }
          if (SSD_X & 8) ml = 0;
          SSD_UNROLL(SSD_F_KU)
          for (int k = 0; k < ml; ++k) {
#pragma unroll
            for (int c = 0; c < CP; ++c) {
              const int idx = ca[c] + k < cb[c] ? ca[c] + k : F_SPAD + 32;  // out of range: a slot that always holds zero
#pragma unroll
              for (int f = 0; f < FR; ++f) acc[c][f] += gfp[f][idx];
            }
          }
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            const int v = lane + 32 * (ci + c);
            float pv[FR];
#pragma unroll
            for (int f = 0; f < FR; ++f) pv[f] = v < V ? pblk[v * F_KS + f0 + r0 + (f < nf ? f : 0)] : 0.f;
#pragma unroll
            for (int f = 0; f < FR; ++f) {
              tot[f] += acc[c][f];
              const float gv = acc[c][f] + (v == P.blank ? gb[f] : 0.f);
              if (!(SSD_X & 16) && f < nf && v < V) st_stream(g_base + (int64_t)(t0 + f0 + r0 + f) * P.st + v, (pv[f] - gv) * scale);
            }
          }
        }
        // D3: the posteriors of a frame sum to one unless mass was lost to fp32 range somewhere
#pragma unroll
        for (int f = 0; f < FR; ++f) {
          const float t = (SSD_X & 32) ? 1.f : warp_sum_fx(tot[f]);
          if (f < nf && !(fabsf(t - 1.f) <= F_MASS_TOL)) mass_ok = false;
        }
      }
      __syncwarp();
    }
    if (stop) break;
    i = i_next;
  }
  if (!mass_ok && lane == 0) atomicOr(&sc->uflag, 8);  // D3 (the sums are warp-uniform)
  if (late) fence_proxy_async();  // my ordinary stores to the ring space precede the next utterance's bulk copies into it
}

// ------------------------------------------------------------------ kernel
// ------------------------------------------------------------------ prepass: utterances -> strip-width classes
struct PrepassParams {
  const int64_t* in_len;
  const int64_t* tg_len;
  int B;
  int64_t T;
  int S_max, nc_max, nc_min;
  int rank;  // 1: order every class by decreasing cost (B <= F_RANK_MAX)
  int* list;
  int* start;
  int* count;
};

// class (0..F_NCLS-1) and cost of utterance j.  Rows the fast path does not take (bad lengths, no frames, a lattice wider
// than the launch holds) ride in the widest class at cost 0: the kernel there flags them for the exact path unseen.
__device__ __forceinline__ int prepass_class(const PrepassParams& P, int j, int& cost) {
  const int64_t tl = P.in_len[j], sl = P.tg_len[j];
  const bool bad = tl < 0 || tl > P.T || sl < 0 || sl > P.S_max || tl == 0 || 2 * sl + 1 > 32 * (int64_t)P.nc_max;
  int need = bad ? P.nc_max : (int)((2 * sl + 1 + 31) / 32);
  need = need < P.nc_min ? P.nc_min : need;
  cost = bad ? 0 : (int)tl;
  return fast_class_of(need);
}

__global__ void __launch_bounds__(1024) loss_fast_prepass_kernel(const PrepassParams P) {
  __shared__ int cnt[F_NCLS], first[F_NCLS], fill[F_NCLS];
  __shared__ unsigned long long key[F_RANK_MAX];
  const int tid = (int)threadIdx.x;
  if (tid < F_NCLS) cnt[tid] = fill[tid] = 0;
  __syncthreads();
  for (int j = tid; j < P.B; j += 1024) {
    int cost;
    const int c = prepass_class(P, j, cost);
    atomicAdd(&cnt[c], 1);
    if (P.rank) key[j] = ((unsigned long long)c << 40) | (unsigned long long)(unsigned)cost;
  }
  __syncthreads();
  if (tid == 0) {
    int o = 0;
    for (int c = F_NCLS - 1; c >= 0; --c) {  // widest class first
      first[c] = o;
      P.start[c] = o;
      P.count[c] = cnt[c];
      o += cnt[c];
    }
  }
  __syncthreads();
  if (P.rank) {
    // (class, cost) descending, index ascending: the rank IS the position in the list
    for (int j = tid; j < P.B; j += 1024) {
      const unsigned long long mine = key[j];
      int rk = 0;
      for (int i = 0; i < P.B; ++i) {
        const unsigned long long o = key[i];
        rk += (o > mine || (o == mine && i < j)) ? 1 : 0;
      }
      P.list[rk] = j;
    }
  } else {
    for (int j = tid; j < P.B; j += 1024) {
      int cost;
      const int c = prepass_class(P, j, cost);
      P.list[first[c] + atomicAdd(&fill[c], 1)] = j;  // any order within the class: results do not depend on it
    }
  }
}

// ------------------------------------------------------------------ kernel (one per strip width)
template <int NC>
__global__ void __launch_bounds__(F_THREADS, 2) loss_fast_kernel(const FastParams P) {
  SSD_DYN_SMEM(smem);
  constexpr int CLS = (NC - 4) / 2;
  static_assert(NC >= 4 && NC <= 16 && NC % 2 == 0, "strip widths 4, 6, .., 16");
  // The kernels of one call are launched back to back with programmatic stream serialisation: each lets its
  // successor start at once (they are independent), so all widths run side by side out of ONE stream.  Only the first
  // one waits for the prepass; its successors cannot start before every one of its CTAs has passed that wait.
#ifndef SSD_EMU
  if (P.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
  const int n_cls = P.direct ? P.B : P.cls_count[CLS];
  const int* my_list = P.direct ? P.order : P.cls_list + P.cls_start[CLS];
  unsigned* my_ctr = P.counter + 16 + 4 * CLS;
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // Who works is decided by the FIRST DRAW, not by the block index: every width's kernel has as many CTAs as there are
  // slots, the block scheduler hands them out from SM 0 upwards, and with "block index < utterances of this width"
  // all widths piled their work onto the same few SMs.  A CTA that finds a working CTA already on its SM lets the
  // CTAs on idle SMs draw first; whoever draws nothing leaves before touching shared memory.
  // my_ctr[0] counts draws (at most n_cls succeed); [1] / [2] are the two ends of the class list (cost-ordered when
  // ranked: the first CTA of an SM draws from the expensive end, a second one from the cheap end, later draws always
  // from the expensive end).  heads + tails <= n_cls, so the ends never cross.
  __shared__ int first_draw[4];  // utterance position, SM id, working CTAs found on the SM
  if (tid == 0) {
#ifdef SSD_EMU
    const unsigned smid = 0u, here = blockIdx.x & 1u;
#else
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned here = *reinterpret_cast<volatile unsigned*>(P.sm_slot + smid);
    if (here > 0u && !P.direct) {
      // a working CTA of the SAME width here: sharing its code in the instruction cache is worth more than an idle SM
      // is to a narrower width; of another width: everybody else first
      const unsigned other = *reinterpret_cast<volatile unsigned*>(P.sm_slot + 256 + smid);
      __nanosleep(other == (unsigned)CLS + 1u ? F_YIELD_NS / 4 : 2 * F_YIELD_NS);
    }
#endif
    int u = n_cls > 0 ? (int)atomicAdd(my_ctr, 1u) : 0;
    if (u < n_cls) {
      const bool tail = P.ranked && here > 0u;
      const int k = (int)atomicAdd(my_ctr + (tail ? 2 : 1), 1u);
      u = tail ? n_cls - 1 - k : k;
    } else {
      u = -1;
    }
    first_draw[0] = u;
    first_draw[1] = (int)smid;
  }
  __syncthreads();
  if (first_draw[0] < 0) return;  // more CTAs than this width has utterances
  const FastSmem M = fast_smem_layout(P.V, P.RS, P.S_max + 4, P.nb_max);

  int* lab = reinterpret_cast<int*>(smem + M.lab_off);
  int* pos = reinterpret_cast<int*>(smem + M.pos_off);
  int* cls = reinterpret_cast<int*>(smem + M.cls_off);
  int* tab = reinterpret_cast<int*>(smem + M.tab_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + M.bar_off);
  uint64_t* blk_ready = reinterpret_cast<uint64_t*>(smem + M.blk_off);
  // The offsets here depend on launch parameters only, so ptxas keeps them in uniform registers -- and ptxas 12.9 has
  // been caught (cuda-gdb + SASS: `MOV R, UR4 ; LDS [R+UR8+0x14]` at the workers' entry with UR4 long since holding
  // V+1) re-materialising this kernel's copy of the scalar block's offset from a uniform register that the role code
  // had reused: every access to *sc then went to a wrong, even misaligned, shared address, in builds that differed
  // from a working one by a moved load.  The offset therefore takes a round trip through shared memory (a volatile
  // load cannot be re-materialised from anything but memory); a shuffle is not enough, it is folded away.
#ifndef SSD_EMU
  __shared__ unsigned pin_scal;
  if (tid == 0) pin_scal = (unsigned)M.scal_off;
#endif
  uint64_t* z_ready = bars + FB_Z;

  if (tid == 0) {
    for (int i = 0; i < FB_TOTAL; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  // One object per block, initialised ONCE per launch (MF's and MB's checkpoint = 2 arrivals).  Every utterance
  // moves every object on by exactly one phase -- the blocks it does not have get their two arrivals right away --
  // so the phase to wait for is the utterance count's parity.  (Re-initialising the objects per utterance left
  // workers asleep on the B200 when a CTA alone on its SM ran through three or more short utterances.)
  for (int q = tid; q < P.nb_max; q += F_THREADS) mbar_init(&blk_ready[q], 2);
  // the dummy class (row V) of every probability block is zero; the workers' stay so for the whole launch, the
  // sweeps' are restored per utterance (a late worker's rows overlay them)
  for (int i = tid; i < (2 * F_NP + F_W) * F_K; i += F_THREADS) {
    const int blk = i / F_K, r = i - blk * F_K;
    unsigned char* base = blk < 2 * F_NP ? smem + M.main_off + (size_t)(blk / F_NP) * M.main_bytes + (size_t)(blk % F_NP) * M.p_slot
                                         : smem + M.worker_off + (size_t)(blk - 2 * F_NP) * M.worker_bytes + M.wrow_bytes;
    reinterpret_cast<float*>(base)[P.V * F_KS + r] = 0.f;
  }
  __syncthreads();
#ifdef SSD_EMU
  FastScal* sc = reinterpret_cast<FastScal*>(smem + M.scal_off);
#else
  FastScal* sc = reinterpret_cast<FastScal*>(smem + *reinterpret_cast<volatile unsigned*>(&pin_scal));
#endif

  unsigned qbase = 0u, uiter = 0u, nraw = 0u;
  // Warp w issues from sub-partition w % 4.  Two CTAs share an SM; the second one to arrive uses a mirrored role
  // map, so that every sub-partition carries ONE main sweep instead of both CTAs' mains sitting on the same two:
  //   8 warps:  rot 0: MF 0  MB 1  LF 2  LB 3  workers 4..7     rot 1: MF 2  MB 3  LF 0  LB 1  workers 4..7
  //   10 warps: rot 0: MF 0  MB 1  LF 8  LB 9  workers 2..7     rot 1: MF 2  MB 3  LF 8  LB 9  workers 0,1,4..7
  if (tid == 0) {
    const unsigned smid = (unsigned)first_draw[1];
#ifdef SSD_EMU
    const unsigned slot = blockIdx.x & 1u;
#else
    const unsigned slot = atomicAdd(P.sm_slot + smid, 1u) & 1u;  // working CTAs only
    if (slot == 0u) *reinterpret_cast<volatile unsigned*>(P.sm_slot + 256 + smid) = (unsigned)CLS + 1u;
#endif
    sc->rot = (int)slot & P.rotate;
    sc->second = (int)slot;
    // checkpoint region: as many exist as CTAs of these kernels can be resident (two per SM) or, if that is fewer,
    // as there are utterances.  Few utterances: first come, first served.  Otherwise the SM's two regions are taken
    // and returned (a CTA of another width may follow on this SM while others still run elsewhere).
    int region;
    if (!P.region_by_sm) {
      region = (int)atomicAdd(P.counter + 2, 1u);
    } else {
      unsigned* busy = P.sm_slot + 512 + 2 * smid;
      for (region = -1; region < 0;) {
        if (atomicCAS(busy, 0u, 1u) == 0u) region = (int)(2 * smid);
        else if (atomicCAS(busy + 1, 0u, 1u) == 0u) region = (int)(2 * smid + 1);
#ifndef SSD_EMU
        else __nanosleep(200);
#endif
      }
    }
    sc->region = region;
  }
  __syncthreads();
  unsigned char* ckpt_cta = P.ckpt + (size_t)sc->region * P.ckpt_cta_bytes;
  int role;  // 0 MF, 1 MB, 2 LF, 3 LB, 4 + w: worker w
  {
    const int m0 = sc->rot ? 2 : 0;  // first main warp
    if (F_W == 4) {
      role = warp < 4 ? (warp ^ m0) : warp;
    } else {
      if (warp == m0) role = 0;
      else if (warp == m0 + 1) role = 1;
      else if (warp == 8) role = 2;
      else if (warp == 9) role = 3;
      else role = 4 + (warp < 4 ? (warp & 1) : warp - 2);  // the two free low warps are workers 0,1; warps 4..7 workers 2..5
    }
  }

  for (;;) {
    if (tid == 0) {
      int u = first_draw[0];  // the draw that admitted this CTA
      first_draw[0] = -1;
      if (u < 0) {
        u = (int)atomicAdd(my_ctr, 1u);
        if (u < n_cls) u = (int)atomicAdd(my_ctr + 1, 1u);
      }
      sc->utt = u < n_cls ? 0 : P.B;
      if (u < n_cls) {
        const int b = my_list ? my_list[u] : u;
        const int64_t tl = P.in_len[b], sl = P.tg_len[b];
        const bool bad = tl < 0 || tl > P.T || sl < 0 || sl > P.S_max;
        sc->b = b;
        sc->T = bad ? 0 : (int)tl;
        sc->S = bad ? 0 : (int)sl;
        // everything that is not a plain utterance goes to the exact path unseen
        sc->skip = (bad || tl == 0 || 2 * sl + 1 > 32 * NC) ? 1 : 0;
        sc->uflag = sc->skip;
        sc->next_blk = 0;
        float mul = P.grad_mul ? *P.grad_mul : 1.f;
        if (P.mean_den > 0) mul /= (float)((double)P.mean_den * (double)(sl > 1 ? sl : 1));
        sc->scale = mul;
      }
    }
    __syncthreads();
    if (sc->utt >= P.B) break;
    const int b = sc->b, S = sc->S;
#if defined(SSD_LOSS_UTTRACE) && !defined(SSD_EMU)
    if (P.tlog && tid == 0) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      P.tlog[4 * b + 0] = t;
      P.tlog[4 * b + 2] = first_draw[1];
      P.tlog[4 * b + 3] = NC;
    }
#endif
    if (P.late_worker && uiter > 0u) {  // the previous utterance's late worker had its rows over the sweeps' rings
      SSD_F_SETUP_ROLL
      for (int i = tid; i < 2 * F_NP * F_K; i += F_THREADS) {
        const int blk = i / F_K, r = i - blk * F_K;
        unsigned char* base = smem + M.main_off + (size_t)(blk / F_NP) * M.main_bytes + (size_t)(blk % F_NP) * M.p_slot;
        reinterpret_cast<float*>(base)[P.V * F_KS + r] = 0.f;
      }
    }
    SSD_F_SETUP_ROLL
    for (int j = tid; j < S; j += F_THREADS) {
      const int64_t y = P.targets[(int64_t)b * P.tgt_stride + j];
      if (y < 0 || y >= P.V) sc->skip = sc->uflag = 1;  // benign race: every writer stores 1
      lab[j] = (int)(y < 0 ? 0 : (y >= P.V ? P.V - 1 : y));
    }
    __syncthreads();
    const bool skip = sc->skip != 0;
    FPlan pl;
    pl.b = b; pl.Tb = sc->T; pl.S = S; pl.L = 2 * S + 1;
    pl.NB = (pl.Tb + F_K - 1) / F_K;
    if (!skip) {
      SSD_F_SETUP_ROLL
      for (int q = pl.NB + tid; q < P.nb_max; q += F_THREADS) {  // blocks this utterance does not have
        mbar_arrive(&blk_ready[q]);
        mbar_arrive(&blk_ready[q]);
      }
      // class histogram -> start offsets; position of every label in class-sorted order (stable)
      SSD_F_SETUP_ROLL
      for (int v = tid; v < P.V; v += F_THREADS) {
        int c = 0;
        SSD_F_SETUP_ROLL
        for (int j = 0; j < S; ++j) c += (lab[j] == v) ? 1 : 0;
        cls[v + 1] = c;
      }
      __syncthreads();
      if (tid == 0) {
        cls[0] = 0;
        SSD_F_SETUP_ROLL
        for (int v = 0; v < P.V; ++v) cls[v + 1] += cls[v];
      }
      __syncthreads();
      SSD_F_SETUP_ROLL
      for (int j = tid; j < S; j += F_THREADS) {
        const int y = lab[j];
        int k = 0;
        SSD_F_SETUP_ROLL
        for (int i = 0; i < j; ++i) k += (lab[i] == y) ? 1 : 0;
        pos[j] = cls[y] + k;
      }
      __syncthreads();
      build_lane_tables<NC>(S, P.V, lab, pos, tab);
      __syncthreads();
      // rows the sweeps never visit: frames beyond the input length
      if (P.grad) {
        float* g0 = P.grad + (int64_t)b * P.sb;
        SSD_F_SETUP_ROLL
        for (int64_t r = pl.Tb + warp; r < P.T; r += F_THREADS / 32)
          SSD_F_SETUP_ROLL
          for (int v = lane; v < P.V; v += 32) st_stream(g0 + r * P.st + v, 0.f);
      }
      int wslot = role >= 4 ? role - 4 : -1;  // worker buffer set: 0..F_W-1, F_W = the sweeps' ring space (late worker)
      if (role < 4) {
        const int h = role & 1;
        MainView mv;
        mv.base = smem + M.main_off + (size_t)h * M.main_bytes;
        mv.bar = bars + h * FB_PER_MAIN;
        mv.p_slot = M.p_slot;
        mv.raw_slot = M.raw_slot;
        if (role == 0) fast_main<NC, 0>(P, mv, pl, tab, qbase, ckpt_cta, blk_ready, sc, uiter);
        else if (role == 1) fast_main<NC, 1>(P, mv, pl, tab, qbase, ckpt_cta, blk_ready, sc, uiter);
        else if (role == 2) fast_loader<0>(P, mv, pl, qbase, sc, uiter);
        else fast_loader<1>(P, mv, pl, qbase, sc, uiter);
        // Both sweeps and both loaders are done well before the workers are (the workers' share is ~2/3 of an
        // utterance's time): one of these four warps joins the workers for the rest of the utterance, with the
        // sweeps' ring space -- idle from here on -- as its buffers.
        if (P.late_worker && P.grad) {
          bar_sync(2, 128);  // all four: nobody reads or fills the rings any more
          if (role == 2) wslot = F_W;
          if (role == 3 && P.late_worker > 1) wslot = F_W + 1;  // room for a second one (smaller row buffers, deeper rings)
        }
      }
      if (wslot >= 0) {
        const bool late = wslot >= F_W;
        unsigned char* wb = late ? smem + M.main_off + (size_t)(wslot - F_W) * M.worker_bytes : smem + M.worker_off + (size_t)wslot * M.worker_bytes;
        WorkerView wv;
        wv.rows = reinterpret_cast<float*>(wb);
        wv.pblk = reinterpret_cast<float*>(wb + M.wrow_bytes);
        wv.raw0 = reinterpret_cast<float*>(wb + M.wrow_bytes + M.p_slot);
        wv.raw_stride = (int)(M.raw_slot / 4);
        wv.rawbar = bars + FB_WRAW + 2 * wslot;
        if (late) {
          if (lane < F_K) wv.pblk[P.V * F_KS + lane] = 0.f;  // the dummy class of this block buffer
          __syncwarp();
          fence_proxy_async();  // bytes last written by ordinary stores are about to be written by bulk copies
        }
        fast_worker<NC>(P, pl, tab, cls, wv, ckpt_cta, blk_ready, z_ready, uiter, nraw, sc, late);  // the only call site
      }
      qbase += (unsigned)pl.NB;
      ++uiter;
    }
    __syncthreads();
    if (tid == 0) {
      const int f = sc->uflag;
      P.flags[b] = f;
      if (f) atomicAdd(P.counter + 1, 1u);
#if defined(SSD_LOSS_UTTRACE) && !defined(SSD_EMU)
      if (P.tlog) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.tlog[4 * b + 1] = t;
      }
#endif
    }
    __syncthreads();  // label tables and scalars are rebuilt for the next utterance
  }
  if (P.region_by_sm && tid == 0) {
    __threadfence();
    atomicExch(P.sm_slot + 512 + sc->region, 0u);
  }
}

// ------------------------------------------------------------------ host side
struct FastPlan {
  int NC, RS, grid, nb_max, regions;
  size_t smem, ckpt_cta_bytes, ckpt_rec_bytes, ws_bytes;
  size_t off_flags, off_sm, off_cls, off_list, off_ckpt, off_tlog;
};

static int fast_env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s && *s ? std::atoi(s) : dflt;
}

static int fast_device_sms() {
#ifdef SSD_EMU
  return 2;
#else
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
#endif
}

static bool fast_plan_for(int64_t B, int64_t T, int64_t V, int64_t S_max, FastPlan& pl) {
  if (fast_env_int("SSD_LOSS_EXACT", 0)) return false;  // force the fp64 path (tests, comparisons)
  if (S_max > 255 || V > 128 || T < 1 || B < 1) return false;
  const int need = (int)((2 * S_max + 1 + 31) / 32);
  if (need > 16) return false;
  pl.NC = fast_width_of(fast_class_of(need));  // the widest strip of the launch: shared memory is laid out for it
  const int cells = 32 * pl.NC, stage = F_SPAD + 32 + 4;
  pl.RS = ((cells > stage ? cells : stage) + 3) & ~3;
  pl.nb_max = (int)((T + F_K - 1) / F_K);
  pl.smem = fast_smem_layout((int)V, pl.RS, (int)S_max + 4, pl.nb_max).total;
  if (pl.smem > 113 * 1024) return false;
  const int64_t slots = (int64_t)fast_device_sms() * 2;
  pl.grid = (int)(B < slots ? B : slots);  // per strip-width kernel; CTAs beyond a width's utterance count leave at once
  pl.regions = pl.grid;
  {  // test knob: fewer CTAs than utterances, so that a CTA walks through several of them
    const int cap = fast_env_int("SSD_LOSS_GRID", 0);
    if (cap > 0 && cap < pl.grid) pl.grid = cap;
  }
  pl.ckpt_rec_bytes = (size_t)pl.NC * 32 * 4 + 128;
  pl.ckpt_cta_bytes = (size_t)2 * pl.nb_max * pl.ckpt_rec_bytes;
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  size_t o = 256;
  pl.off_flags = o; o += al((size_t)B * 4);
  pl.off_sm = o; o += 4096;
  pl.off_cls = o; o += 256;
  pl.off_list = o; o += al((size_t)B * 4);
  pl.off_ckpt = o; o += (size_t)pl.regions * pl.ckpt_cta_bytes;
  pl.off_tlog = o; o += (size_t)(F_THREADS / 32) * F_TLOG_N * F_TLOG_S * 8;
  pl.ws_bytes = o;
  return true;
}

size_t loss_fast_workspace_bytes(int64_t B, int64_t T, int64_t V, int64_t S_max) {
  FastPlan pl;
  return fast_plan_for(B, T, V, S_max, pl) ? pl.ws_bytes : 0;
}

template <int NC>
static int loss_fast_launch_t(const FastParams& P, const FastPlan& pl, cudaStream_t stream) {
  auto kfn = loss_fast_kernel<NC>;
#ifndef SSD_EMU
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "loss(fast): cudaFuncSetAttribute(%zu): %s", pl.smem, cudaGetErrorString(e));
  // programmatic stream serialisation: this kernel may start as soon as its predecessor in the stream lets it
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)pl.grid);
  cfg.blockDim = dim3(F_THREADS);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kfn, P);
  if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "loss_fast_kernel<%d>: %s", NC, cudaGetErrorString(e));
#else
  SSD_LAUNCH(kfn, (unsigned)pl.grid, F_THREADS, pl.smem, stream, P);
#endif
  return check_launch("loss_fast_kernel");
}

// Returns SSD_OK and sets *flags / *nflag (device pointers into ws) when the fast path ran,
// SSD_ERR_UNSUPPORTED (without touching the error text) when it does not apply.
int loss_fast_launch(const float* lp, const float* lse, int64_t lse_stride, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st,
                     const int64_t* targets, int64_t tgt_stride, int64_t S_max, const int64_t* in_len, const int64_t* tg_len,
                     const int32_t* order, int blank, int zero_inf, int64_t mean_den, const float* grad_mul, float* nll, float* grad,
                     void* ws, size_t ws_bytes, const int** flags, const unsigned** nflag, cudaStream_t stream) {
  FastPlan pl;
  if (!fast_plan_for(B, T, V, S_max, pl)) return SSD_ERR_UNSUPPORTED;
  if (!ws || ws_bytes < pl.ws_bytes) return fail(SSD_ERR_WORKSPACE, "ssd_ctc_loss: workspace %zu bytes, need %zu", ws_bytes, pl.ws_bytes);
  // `order` is honoured by the direct scheme only; otherwise the prepass groups and orders the utterances on the device
  // (by strip width, then by frames)
  FastParams P;
  P.lp = lp; P.lse = lse; P.lse_stride = lse_stride; P.T = T; P.sb = sb; P.st = st; P.B = (int)B; P.V = (int)V;
  P.targets = targets; P.tgt_stride = tgt_stride; P.S_max = (int)S_max;
  P.in_len = in_len; P.tg_len = tg_len; P.order = order;
  P.blank = blank; P.zero_inf = zero_inf; P.mean_den = mean_den; P.grad_mul = grad_mul;
  P.nll = nll; P.grad = grad;
  unsigned char* w = reinterpret_cast<unsigned char*>(ws);
  P.counter = reinterpret_cast<unsigned*>(w);
  P.flags = reinterpret_cast<int*>(w + pl.off_flags);
  P.sm_slot = reinterpret_cast<unsigned*>(w + pl.off_sm);
  int* cls = reinterpret_cast<int*>(w + pl.off_cls);
  P.cls_start = cls;
  P.cls_count = cls + 8;
  P.cls_list = reinterpret_cast<int*>(w + pl.off_list);
  P.ckpt = w + pl.off_ckpt;
  P.ckpt_cta_bytes = pl.ckpt_cta_bytes; P.ckpt_rec_bytes = pl.ckpt_rec_bytes; P.nb_max = pl.nb_max;
  P.RS = pl.RS;
  P.use_tma = (st == V) && aligned16(lp) && ((sb * 4) % 16 == 0) && ((V * 4) % 16 == 0) ? 1 : 0;
  if (lse && !(aligned16(lse) && lse_stride % F_K == 0)) P.use_tma = 0;
  if (fast_env_int("SSD_LOSS_NOTMA", 0)) P.use_tma = 0;
  P.rotate = fast_env_int("SSD_LOSS_NOROT", 0) ? 0 : 1;
  {
    const FastSmem m = fast_smem_layout((int)V, pl.RS, (int)S_max + 4, pl.nb_max);
    P.late_worker = fast_env_int("SSD_LOSS_NOLATE", 0) ? 0 : (int)(2 * m.main_bytes / m.worker_bytes);
    if (P.late_worker > 2) P.late_worker = 2;
  }
  P.tlog = (fast_env_int("SSD_LOSS_DBG", 0) & 8) ? reinterpret_cast<long long*>(w + pl.off_tlog) : nullptr;
  P.region_by_sm = B > pl.regions ? 1 : 0;
  *flags = P.flags;
  *nflag = P.counter + 1;

  PrepassParams Q;
  Q.in_len = in_len; Q.tg_len = tg_len; Q.B = (int)B; Q.T = T; Q.S_max = (int)S_max; Q.nc_max = pl.NC;
  // Per-utterance strip widths pay when the SMs are busy for several rounds (measured on B200, T=1000, S~U[50,200]:
  // B=1024 -11 %, B=4096 -24 %).  With at most two utterances per SM the launch lasts as long as its slowest pair, whose
  // time hardly depends on the width (sweeps are ~25 % of the instructions), and CTAs of DIFFERENT kernels sharing an
  // SM run slower than two of the same one (instruction cache): B=256 took 432 us against 342 us.  So small batches
  // run every utterance at the launch-wide width, in one kernel.
  Q.nc_min = B > 4 * (int64_t)fast_device_sms() ? 4 : pl.NC;
  {  // tuning / test knobs: force one or the other scheme, or a minimum strip width
    const int multi = fast_env_int("SSD_LOSS_MULTI", -1);
    if (multi == 1) Q.nc_min = 4;
    if (multi == 0 || fast_env_int("SSD_LOSS_ONEWIDTH", 0)) Q.nc_min = pl.NC;
    const int nc = fast_env_int("SSD_LOSS_NC", 0);
    if (nc >= 4 && nc <= pl.NC) Q.nc_min = fast_width_of(fast_class_of(nc));
  }
  Q.rank = B <= F_RANK_MAX ? 1 : 0;
  Q.list = const_cast<int*>(P.cls_list); Q.start = cls; Q.count = cls + 8;
  P.ranked = (Q.rank && !fast_env_int("SSD_LOSS_NODEAL", 0)) ? 1 : 0;
#ifndef SSD_EMU
  cudaError_t e = cudaMemsetAsync(P.counter, 0, 256, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(P.sm_slot, 0, 4096, stream);
  if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "loss(fast): cudaMemsetAsync: %s", cudaGetErrorString(e));
#else
  for (int i = 0; i < 64; ++i) P.counter[i] = 0u;
  for (int i = 0; i < 1024; ++i) P.sm_slot[i] = 0u;
#endif
  // widest first: its utterances are the longest-running ones
  const int c_hi = fast_class_of(pl.NC), c_lo = fast_class_of(Q.nc_min);
  // one width and a checkpoint region for every utterance: each CTA takes one utterance, nothing to group or rank
  P.direct = (c_lo == c_hi && B <= pl.regions && pl.grid == pl.regions && !fast_env_int("SSD_LOSS_PREPASS", 0)) ? 1 : 0;
  int rc = SSD_OK;
  if (!P.direct) {
    SSD_LAUNCH(loss_fast_prepass_kernel, 1u, 1024, 0, stream, Q);
    rc = check_launch("loss_fast_prepass_kernel");
    if (rc) return rc;
  } else {
    P.ranked = 0;
  }
  for (int c = c_hi; c >= c_lo; --c) {
    P.pdl_wait = (c == c_hi && !P.direct) ? 1 : 0;
    switch (c) {
      case 0: rc = loss_fast_launch_t<4>(P, pl, stream); break;
      case 1: rc = loss_fast_launch_t<6>(P, pl, stream); break;
      case 2: rc = loss_fast_launch_t<8>(P, pl, stream); break;
      case 3: rc = loss_fast_launch_t<10>(P, pl, stream); break;
      case 4: rc = loss_fast_launch_t<12>(P, pl, stream); break;
      case 5: rc = loss_fast_launch_t<14>(P, pl, stream); break;
      default: rc = loss_fast_launch_t<16>(P, pl, stream); break;
    }
    if (rc) return rc;
  }
  return SSD_OK;
}

}  // namespace ssd
