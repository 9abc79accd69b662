// CTC loss, forward + backward in ONE launch: nll[B] and d nll / d log_probs.
//
// Replaces torch.nn.CTCLoss as called at reference src/models/losses.py:56 (ATen ctc_loss + its
// autograd backward).  See DESIGN.md section 4 for the derivation and the roofline discussion.
//
// Numerics.  The alpha/beta recursions run in the LINEAR domain on scaled fp64 values
//   x_t[s] = (x_{t-1}[s] + x_{t-1}[s-1] + skip[s] * x_{t-1}[s-2]) * p_t[l'_s]
// with an exact power-of-two rescale every few frames and the exponent tracked as an integer.
// No exp/log sits on the recursion (ATen's log-space form needs 3-4 MUFU ops per cell) and the
// result is accurate to ~1e-13, whereas fp32 log-space loses ~1e-3 of the gradient at T=1000.
// p = expf(lp) is evaluated once per (frame, label) in fp32 and widened.
//
// Parallel structure (one CTA = 4 warps per utterance, CTAs pull utterances from a queue):
//   warp 0  F-main   forward sweep  (alpha), local time tau = t
//   warp 1  R-main   backward sweep (beta),  local time tau = T-1-t  -- by the symmetry of CTC this
//                    is the same recursion on the reversed label sequence
//   warp 2/3 helpers, one per main: producer side of a two-stage mbarrier pipeline.
// A lane keeps NC consecutive lattice cells in registers; one shuffle per frame moves the strip
// boundary.  The two sweeps MEET IN THE MIDDLE: each main first runs half of the frames
// (phase 1), checkpointing its strip every K frames to an L2-resident workspace, then continues
// through the other half (phase 2) where it needs the other sweep's values frame by frame.
// Those are re-computed K frames at a time from the checkpoints by the main's helper (which
// therefore runs the *other* direction's recursion) into shared memory, so alpha and beta are
// never materialised in HBM: traffic is log_probs once + gradient once + (1/K) checkpoint spill.
// In phase 2 a main multiplies its fresh x by the helper's pre-emission sums to get the lattice
// posteriors, reduces them per label class (blank: warp shuffle; others: label-sorted staging in
// shared memory) and writes the gradient row  scale * (p - posterior)  straight to HBM.
//
// Log-prob rows reach shared memory through 1-D TMA bulk copies (cp.async.bulk + mbarrier,
// SASS UBLKCP) issued 3 segments ahead by the helpers, which also do the exp() conversion.
#include <cmath>
#include <cstdlib>

#include "ssd_common.cuh"

namespace ssd {

constexpr int LS_THREADS = 128;
constexpr int LS_NRAW = 4;     // raw-row ring depth per half (TMA prefetch distance = 3 segments)
constexpr int LS_RESCALE = 4;  // frames between rescale checks

// ------------------------------------------------------------------ scalar helpers
template <typename R>
struct Fp;
template <>
struct Fp<double> {
  static constexpr int WINDOW = 256;
  static __device__ __forceinline__ int hi(double v) { return __double2hiint(v); }
  static __device__ __forceinline__ int expo_of_hi(int h) { return ((h >> 20) & 0x7ff) - 1023; }
  static __device__ __forceinline__ int clamp_shift(int s) { return s < -1000 ? -1000 : (s > 1000 ? 1000 : s); }
  static __device__ __forceinline__ double pow2(int e) { return __hiloint2double((e + 1023) << 20, 0); }
};
template <>
struct Fp<float> {
  static constexpr int WINDOW = 32;
  static __device__ __forceinline__ int hi(float v) { return __float_as_int(v); }
  static __device__ __forceinline__ int expo_of_hi(int h) { return ((h >> 23) & 0xff) - 127; }
  static __device__ __forceinline__ int clamp_shift(int s) { return s < -120 ? -120 : (s > 120 ? 120 : s); }
  static __device__ __forceinline__ float pow2(int e) { return __int_as_float((e + 127) << 23); }
};
// 2^e for any int e as a double (used for the per-frame posterior scale; may under/overflow)
__device__ __forceinline__ double pow2_wide(int e) {
  if (e < -1070) return 0.0;
  if (e > 1023) return INFINITY;
  if (e >= -1022) return __hiloint2double((e + 1023) << 20, 0);
  return __hiloint2double((e + 1023 + 60) << 20, 0) * __hiloint2double((1023 - 60) << 20, 0);
}

struct LossParams {
  const float* lp;
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
  size_t ckpt_cta_bytes;
  int nseg_max;
  int K;       // frames per segment / pipeline stage
  int NCmax;   // cells per lane the shared-memory stages are sized for
  int VP;      // p-row stride in elements (V + dummy zero column, even)
  int Spad;    // staging capacity (labels)
  int use_tma;
};

struct LossSmem {  // offsets into dynamic shared memory, identical on host and device
  size_t stage_bytes, pR_off, wR_off, wE_off;  // within a stage
  size_t stages_off, raw_off, raw_slot_bytes, gst_off, lab_off, pos_off, cls_off, bar_off, scal_off, total;
};

template <typename R>
__host__ __device__ inline LossSmem loss_smem_layout(int K, int NCmax, int V, int VP, int Spad) {
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  LossSmem m;
  m.pR_off = 0;
  m.wR_off = al((size_t)K * VP * sizeof(R));
  m.wE_off = m.wR_off + al((size_t)K * NCmax * 32 * sizeof(R));
  m.stage_bytes = m.wE_off + al((size_t)K * 4);
  size_t o = 0;
  m.stages_off = o; o += 4 * m.stage_bytes;                     // [half][stage]
  m.raw_slot_bytes = al((size_t)K * V * 4);
  m.raw_off = o; o += 2 * LS_NRAW * m.raw_slot_bytes;           // [half][slot]
  m.gst_off = o; o += 2 * al((size_t)Spad * 4);                 // [half]
  m.lab_off = o; o += al((size_t)Spad * 4);
  m.pos_off = o; o += al((size_t)Spad * 4);
  m.cls_off = o; o += al((size_t)(V + 2) * 4);
  m.bar_off = o; o += al((size_t)(4 + 4 + 2 * LS_NRAW + 2) * 8);  // full[2][2] empty[2][2] raw[2][NRAW] p1done[2]
  m.scal_off = o; o += 64;
  m.total = o;
  return m;
}

struct LossScal {
  int utt, b, T, S, bad, nc;
  float scale;
};

// ------------------------------------------------------------------ the recursion step
// x: strip of NC cells (cell index = lane*NC + i, even i = blank).  On return w holds the
// pre-emission sums and x the post-emission values of the new frame.
template <typename R, int NC>
__device__ __forceinline__ void ctc_step(R (&x)[NC], R (&w)[NC], const int (&lam)[NC / 2], unsigned skipm,
                                         const R* __restrict__ prow, int blank) {
  R up = __shfl_up_sync(FULL, x[NC - 1], 1);
  if (lane_id() == 0) up = R(0);
  const R pb = prow[blank];
#pragma unroll
  for (int i = NC - 1; i >= 0; --i) {
    const R a1 = (i >= 1) ? x[i >= 1 ? i - 1 : 0] : up;
    R s = x[i] + a1;
    if (i & 1) {
      const R a2 = (i >= 2) ? x[i >= 2 ? i - 2 : 0] : up;
      if ((skipm >> (i >> 1)) & 1u) s += a2;
      w[i] = s;
      x[i] = s * prow[lam[i >> 1]];
    } else {
      w[i] = s;
      x[i] = s * pb;
    }
  }
}

template <typename R, int NC>
__device__ __forceinline__ void ctc_rescale(R (&x)[NC], int& E) {
  int m = 0;
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int h = Fp<R>::hi(x[i]);
    m = h > m ? h : m;
  }
  m = __reduce_max_sync(FULL, m);
  if (m == 0) return;
  const int e = Fp<R>::expo_of_hi(m);
  if (e > Fp<R>::WINDOW || e < -Fp<R>::WINDOW) {
    const int sh = Fp<R>::clamp_shift(-e);
    const R f = Fp<R>::pow2(sh);
#pragma unroll
    for (int i = 0; i < NC; ++i) x[i] *= f;
    E -= sh;
  }
}

// per-lane label columns / skip mask for sweep direction dir (0 forward, 1 reversed labels)
template <int NC>
__device__ __forceinline__ void lane_labels(int dir, int S, int V, const int* __restrict__ lab, int (&lam)[NC / 2], unsigned& skipm) {
  const int lane = lane_id();
  skipm = 0u;
#pragma unroll
  for (int k = 0; k < NC / 2; ++k) {
    const int jl = lane * (NC / 2) + k;  // local label index of odd cell 2k+1
    int col = V;                         // dummy column (p = 0) beyond the target
    if (jl < S) {
      const int j = dir ? S - 1 - jl : jl;
      col = lab[j];
      if (jl >= 1) {
        const int jp = dir ? j + 1 : j - 1;
        if (lab[jp] != col) skipm |= 1u << k;
      }
    }
    lam[k] = col;
  }
}

struct Pipe {  // per half
  unsigned char* stage[2];
  uint64_t* full;   // [2]
  uint64_t* empty;  // [2]
  uint64_t* rawbar; // [NRAW]
  float* raw;       // [NRAW][K*V]
  size_t raw_slot_floats;
};

// production descriptor: which K-frame block of which sweep direction
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

// ------------------------------------------------------------------ helper warp
template <typename R, int NC>
__device__ void helper_role(const LossParams& P, const LossSmem& M, const Pipe& pipe, int h, int b, int Tb, int S, int m_h, int m_o,
                            const int* __restrict__ lab, uint64_t* p1done_other, unsigned uiter, unsigned& qg,
                            unsigned char* ckpt_cta) {
  const int lane = lane_id();
  const int K = P.K, V = P.V, VP = P.VP;
  const int L = 2 * S + 1;
  const int NP1 = (m_h + K - 1) / K, NSEG = (m_o + K - 1) / K, NPROD = NP1 + NSEG;
  const int d = h ^ 1;  // direction this helper recomputes
  const float* base = P.lp + (int64_t)b * P.sb;

  int lam[NC / 2];
  unsigned skipm;
  lane_labels<NC>(d, S, V, lab, lam, skipm);
  int woff[NC];  // where my cell's value goes in the main's (reversed) layout, -1 = nowhere
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int sig_o = lane * NC + i;
    int off = -1;
    if (sig_o < L) {
      const int sig = L - 1 - sig_o;
      off = (sig % NC) * 32 + sig / NC;
    }
    woff[i] = off;
  }

  const unsigned rawbase = qg;
  auto issue_raw = [&](int q) {  // lane 0
    const Prod pd = prod_desc(q, h, m_h, m_o, K, NP1, NSEG);
    const int t_lo = pd.dir ? Tb - pd.tau0 - pd.n : pd.tau0;
    const unsigned rq = rawbase + (unsigned)q;
    const int slot = (int)(rq % LS_NRAW);
    const uint32_t bytes = (uint32_t)(pd.n * V * 4);
    mbar_arrive_expect_tx(&pipe.rawbar[slot], bytes);
    bulk_g2s(pipe.raw + slot * pipe.raw_slot_floats, base + (int64_t)t_lo * P.st, bytes, &pipe.rawbar[slot]);
  };
  if (P.use_tma && lane == 0)
    for (int q = 0; q < NPROD && q < LS_NRAW - 1; ++q) issue_raw(q);

  R x[NC], w[NC];
  for (int q = 0; q < NPROD; ++q) {
    if (P.use_tma && lane == 0 && q + LS_NRAW - 1 < NPROD) issue_raw(q + LS_NRAW - 1);
    const int s = (int)(qg & 1u);
    mbar_wait(&pipe.empty[s], ((qg >> 1) & 1u) ^ 1u);
    const Prod pd = prod_desc(q, h, m_h, m_o, K, NP1, NSEG);
    R* pR = reinterpret_cast<R*>(pipe.stage[s] + M.pR_off);
    // ---- rows: exp() into the stage, row r <-> local time tau0 + r of direction pd.dir
    if (P.use_tma) {
      const unsigned rq = rawbase + (unsigned)q;
      const int slot = (int)(rq % LS_NRAW);
      mbar_wait(&pipe.rawbar[slot], (rq / LS_NRAW) & 1u);
      const float* raw = pipe.raw + slot * pipe.raw_slot_floats;
      for (int idx = lane; idx < pd.n * V; idx += 32) {
        const int r = idx / V, v = idx - r * V;
        const int br = pd.dir ? pd.n - 1 - r : r;
        pR[r * VP + v] = (R)expf(raw[br * V + v]);
      }
    } else {
      const int t_lo = pd.dir ? Tb - pd.tau0 - pd.n : pd.tau0;
      for (int idx = lane; idx < pd.n * V; idx += 32) {
        const int r = idx / V, v = idx - r * V;
        const int br = pd.dir ? pd.n - 1 - r : r;
        pR[r * VP + v] = (R)expf(ld_stream(base + (int64_t)(t_lo + br) * P.st + v));
      }
    }
    if (lane < pd.n) pR[lane * VP + V] = R(0);  // dummy column (K <= 32)
    __syncwarp();
    // ---- phase 2: recompute segment of direction d from its checkpoint
    if (q >= NP1) {
      if (q == NP1) mbar_wait(p1done_other, uiter & 1u);  // the other main has published its checkpoints
      const int seg = pd.tau0 / K;
      int E = 0;
      if (seg == 0) {
#pragma unroll
        for (int i = 0; i < NC; ++i) x[i] = R(0);
        if (lane == 0) x[0] = R(1);
      } else {
        const unsigned char* cp = ckpt_cta + ((size_t)d * P.nseg_max + seg) * ((size_t)P.NCmax * 32 * sizeof(R) + 16);
        const R* cx = reinterpret_cast<const R*>(cp);
#pragma unroll
        for (int i = 0; i < NC; ++i) x[i] = cx[i * 32 + lane];
        E = *reinterpret_cast<const int*>(cp + (size_t)P.NCmax * 32 * sizeof(R));
      }
      R* wR = reinterpret_cast<R*>(pipe.stage[s] + M.wR_off);
      int* wE = reinterpret_cast<int*>(pipe.stage[s] + M.wE_off);
      int tau = pd.tau0;
      for (int r = 0; r < pd.n; ++r, ++tau) {
        const int Eb = E;
        ctc_step<R, NC>(x, w, lam, skipm, pR + r * VP, P.blank);
        R* dst = wR + (size_t)r * NC * 32;
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (woff[i] >= 0) dst[woff[i]] = w[i];
        if (lane == 0) wE[r] = Eb;
        if ((tau & (LS_RESCALE - 1)) == LS_RESCALE - 1) ctc_rescale<R, NC>(x, E);
      }
    }
    mbar_arrive(&pipe.full[s]);  // all 32 lanes: their shared-memory writes are ordered before it
    ++qg;
  }
}

// ------------------------------------------------------------------ main warp
template <typename R, int NC>
__device__ void main_role(const LossParams& P, const LossSmem& M, const Pipe& pipe, int h, int b, int Tb, int S, int m_h, int m_o,
                          const int* __restrict__ lab, const int* __restrict__ pos, const int* __restrict__ cls, float* gst,
                          uint64_t* p1done_mine, float scale, unsigned& qg, unsigned char* ckpt_cta) {
  const int lane = lane_id();
  const int K = P.K, V = P.V, VP = P.VP;
  const int L = 2 * S + 1;
  const int NP1 = (m_h + K - 1) / K, NSEG = (m_o + K - 1) / K;

  int lam[NC / 2], spos[NC / 2];
  unsigned skipm;
  lane_labels<NC>(h, S, V, lab, lam, skipm);
#pragma unroll
  for (int k = 0; k < NC / 2; ++k) {
    const int jl = lane * (NC / 2) + k;
    spos[k] = jl < S ? pos[h ? S - 1 - jl : jl] : -1;
  }
  unsigned validm = 0u;  // cells of mine inside the lattice
#pragma unroll
  for (int i = 0; i < NC; ++i)
    if (lane * NC + i < L) validm |= 1u << i;

  R x[NC], w[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) x[i] = R(0);
  if (lane == 0) x[0] = R(1);
  int E = 0, tau = 0;

  // ---------------- phase 1: my half of the frames, checkpoint every K
  for (int q = 0; q < NP1; ++q) {
    const int s = (int)(qg & 1u);
    mbar_wait(&pipe.full[s], (qg >> 1) & 1u);
    const R* pR = reinterpret_cast<const R*>(pipe.stage[s] + M.pR_off);
    const int n = min(K, m_h - q * K);
    for (int r = 0; r < n; ++r, ++tau) {
      ctc_step<R, NC>(x, w, lam, skipm, pR + r * VP, P.blank);
      if ((tau & (LS_RESCALE - 1)) == LS_RESCALE - 1) ctc_rescale<R, NC>(x, E);
    }
    mbar_arrive(&pipe.empty[s]);
    ++qg;
    if (q + 1 < NP1) {  // state after frame (q+1)K-1 starts segment q+1
      unsigned char* cp = ckpt_cta + ((size_t)h * P.nseg_max + (q + 1)) * ((size_t)P.NCmax * 32 * sizeof(R) + 16);
      R* cx = reinterpret_cast<R*>(cp);
#pragma unroll
      for (int i = 0; i < NC; ++i) cx[i * 32 + lane] = x[i];
      if (lane == 0) *reinterpret_cast<int*>(cp + (size_t)P.NCmax * 32 * sizeof(R)) = E;
    }
  }
  __threadfence_block();
  mbar_arrive(p1done_mine);

  // ---------------- phase 2: the other half, with the other sweep's values from my helper
  bool first = true, feasible = true;
  double invPm = 1.0;
  int eP = 0;
  float* grow_base = P.grad ? P.grad + (int64_t)b * P.sb : nullptr;
  for (int j = 0; j < NSEG; ++j) {
    const int seg = NSEG - 1 - j;
    const int n = min(K, m_o - seg * K);
    const int s = (int)(qg & 1u);
    mbar_wait(&pipe.full[s], (qg >> 1) & 1u);
    const R* pR = reinterpret_cast<const R*>(pipe.stage[s] + M.pR_off);
    const R* wR = reinterpret_cast<const R*>(pipe.stage[s] + M.wR_off);
    const int* wE = reinterpret_cast<const int*>(pipe.stage[s] + M.wE_off);
    for (int r = n - 1; r >= 0; --r, ++tau) {
      const R* prow = pR + r * VP;
      ctc_step<R, NC>(x, w, lam, skipm, prow, P.blank);
      // lattice occupancy: my post-emission value x other's pre-emission sum.  The two factors can
      // each sit hundreds of binades below their sweep's maximum while their product is what
      // matters, so the exponent k = E + Eo - eP is split over both operands before multiplying.
      const R* wo = wR + (size_t)r * NC * 32 + lane;
      const int Eo = wE[r];
      if (first) {
        first = false;
        // largest product exponent over the lattice, and the x-exponent of a cell attaining it
        int key = 0;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
          const double xv = (double)x[i], wv = ((validm >> i) & 1u) ? (double)wo[i * 32] : 0.0;
          if (xv > 0.0 && wv > 0.0) {
            const int ex = Fp<double>::expo_of_hi(__double2hiint(xv)), ew = Fp<double>::expo_of_hi(__double2hiint(wv));
            const int kk = ((ex + ew + 4096) << 12) | (ex + 2048);
            key = kk > key ? kk : key;
          }
        }
        key = __reduce_max_sync(FULL, key);
        feasible = key != 0;
        if (feasible) {
          const int esum = (key >> 12) - 4096, ex = (key & 0xfff) - 2048;
          const double fa = pow2_wide(-ex), fb = pow2_wide(-(esum - ex));
          double ps = 0.0;
#pragma unroll
          for (int i = 0; i < NC; ++i)
            if ((validm >> i) & 1u) ps += ((double)x[i] * fa) * ((double)wo[i * 32] * fb);
#pragma unroll
          for (int o = 16; o; o >>= 1) ps += __shfl_xor_sync(FULL, ps, o);
          const int e2 = Fp<double>::expo_of_hi(__double2hiint(ps));
          const double pm = ps * pow2_wide(-e2);  // in [1,2)
          invPm = 1.0 / pm;
          eP = e2 + esum + E + Eo;
          if (h == 0 && lane == 0) P.nll[b] = (float)(-(log(pm) + (double)eP * 0.69314718055994530942));
        } else if (h == 0 && lane == 0) {
          P.nll[b] = P.zero_inf ? 0.f : INFINITY;
        }
      }
      if (grow_base && feasible) {
        int k = E + Eo - eP;
        k = k < -2000 ? -2000 : (k > 2000 ? 2000 : k);
        const int k1 = k / 2;
        const double f1 = pow2_wide(k1) * invPm, f2 = pow2_wide(k - k1);
#pragma unroll
        for (int i = 0; i < NC; ++i)
          w[i] = ((validm >> i) & 1u) ? (R)(((double)x[i] * f1) * ((double)wo[i * 32] * f2)) : R(0);
      }
      if (grow_base) {
        const int t = h ? Tb - 1 - tau : tau;
        float* grow = grow_base + (int64_t)t * P.st;
        if (feasible) {
          // blank class: even cells (w now holds normalised posteriors)
          double gb = 0.0;
#pragma unroll
          for (int i = 0; i < NC; i += 2) gb += (double)w[i];
          float gbf = (float)gb;
#pragma unroll
          for (int o = 16; o; o >>= 1) gbf += __shfl_xor_sync(FULL, gbf, o);
          // label classes: stage posteriors in label-sorted order, then each lane sums its classes
#pragma unroll
          for (int k = 0; k < NC / 2; ++k)
            if (spos[k] >= 0) gst[spos[k]] = (float)w[2 * k + 1];
          __syncwarp();
          for (int v = lane; v < V; v += 32) {
            float g = 0.f;
            if (v == P.blank) {
              g = gbf;
            } else {
              const int a = cls[v], e = cls[v + 1];
              for (int z = a; z < e; ++z) g += gst[z];
            }
            st_stream(grow + v, ((float)prow[v] - g) * scale);
          }
          __syncwarp();
        } else {
          const float fill = P.zero_inf ? 0.f : NAN;
          for (int v = lane; v < V; v += 32) st_stream(grow + v, fill);
        }
      }
      if ((tau & (LS_RESCALE - 1)) == LS_RESCALE - 1) ctc_rescale<R, NC>(x, E);
    }
    mbar_arrive(&pipe.empty[s]);
    ++qg;
  }
}

// ------------------------------------------------------------------ kernel
template <typename R>
__global__ void __launch_bounds__(LS_THREADS, 1) loss_kernel(const LossParams P) {
  SSD_DYN_SMEM(smem);
  const LossSmem M = loss_smem_layout<R>(P.K, P.NCmax, P.V, P.VP, P.Spad);
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int h = warp & 1;
  const bool helper = warp >= 2;

  int* lab = reinterpret_cast<int*>(smem + M.lab_off);
  int* pos = reinterpret_cast<int*>(smem + M.pos_off);
  int* cls = reinterpret_cast<int*>(smem + M.cls_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + M.bar_off);
  LossScal* sc = reinterpret_cast<LossScal*>(smem + M.scal_off);
  uint64_t* p1done = bars + 8 + 2 * LS_NRAW;

  Pipe pipe;
  pipe.stage[0] = smem + M.stages_off + (size_t)(2 * h + 0) * M.stage_bytes;
  pipe.stage[1] = smem + M.stages_off + (size_t)(2 * h + 1) * M.stage_bytes;
  pipe.full = bars + 2 * h;
  pipe.empty = bars + 4 + 2 * h;
  pipe.rawbar = bars + 8 + LS_NRAW * h;
  pipe.raw_slot_floats = M.raw_slot_bytes / 4;
  pipe.raw = reinterpret_cast<float*>(smem + M.raw_off) + (size_t)h * LS_NRAW * pipe.raw_slot_floats;
  float* gst = reinterpret_cast<float*>(smem + M.gst_off) + (size_t)h * (((size_t)P.Spad * 4 + 15) / 16 * 4);

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 32);               // full / empty: one arrival per lane
    for (int i = 0; i < 2 * LS_NRAW; ++i) mbar_init(&bars[8 + i], 1);  // raw: lane 0 + tx bytes
    mbar_init(&p1done[0], 32);
    mbar_init(&p1done[1], 32);
    mbar_fence_init();
  }
  __syncthreads();

  unsigned char* ckpt_cta = P.ckpt + (size_t)blockIdx.x * P.ckpt_cta_bytes;
  unsigned qg = 0u;  // productions so far on my half's pipeline (same value in main and helper)
  unsigned uiter = 0u;

  for (;; ++uiter) {
    // ---------------- next utterance + its label tables
    if (tid == 0) {
      const int u = (int)atomicAdd(P.counter, 1u);
      sc->utt = u;
      if (u < P.B) {
        const int b = P.order ? P.order[u] : u;
        int64_t tl = P.in_len[b], sl = P.tg_len[b];
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
      const int64_t n = (int64_t)(P.T - Tb) * P.V;
      for (int64_t i = tid; i < n; i += LS_THREADS) {
        const int64_t r = i / P.V, v = i - r * P.V;
        st_stream(g0 + (Tb + r) * P.st + v, fill);
      }
    }
    if (Tb == 0 && tid == 0) P.nll[b] = bad ? NAN : (S == 0 ? 0.f : (P.zero_inf ? 0.f : INFINITY));

    const int m = Tb / 2;
    const int m_h = h == 0 ? m : Tb - m;
    const int m_o = Tb - m_h;
    const int nc_need = (2 * S + 1 + 31) / 32;
#define SSD_ROLE(NCV)                                                                                                          \
  do {                                                                                                                         \
    if (helper)                                                                                                                \
      helper_role<R, NCV>(P, M, pipe, h, b, Tb, S, m_h, m_o, lab, &p1done[h ^ 1], uiter, qg, ckpt_cta);                         \
    else                                                                                                                       \
      main_role<R, NCV>(P, M, pipe, h, b, Tb, S, m_h, m_o, lab, pos, cls, gst, &p1done[h], scale, qg, ckpt_cta);                \
  } while (0)
    if (nc_need <= 2) SSD_ROLE(2);
    else if (nc_need <= 4) SSD_ROLE(4);
    else if (nc_need <= 8) SSD_ROLE(8);
    else if (nc_need <= 14) SSD_ROLE(14);
    else if (nc_need <= 20) SSD_ROLE(20);
    else SSD_ROLE(32);
#undef SSD_ROLE
    __syncthreads();  // label tables and staging are reused by the next utterance
  }
}

// ------------------------------------------------------------------ host side
struct LossPlan {
  int K, NCmax, VP, Spad, ctas_per_sm, grid, nseg_max;
  size_t smem, ckpt_cta_bytes, ws_bytes;
  bool use_float;
};

static int nc_bucket(int need) {
  const int b[6] = {2, 4, 8, 14, 20, 32};
  for (int i = 0; i < 6; ++i)
    if (need <= b[i]) return b[i];
  return -1;
}

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

template <typename R>
static bool plan_for(int64_t B, int64_t T, int64_t V, int64_t S_max, LossPlan& pl) {
  pl.NCmax = nc_bucket((int)((2 * S_max + 1 + 31) / 32));
  if (pl.NCmax < 0) return false;
  pl.VP = (int)((V + 2) & ~1LL);
  pl.Spad = (int)((S_max + 3) / 4 * 4 + 4);
  const size_t cap = 227 * 1024;
  auto bytes = [&](int K) { return loss_smem_layout<R>(K, pl.NCmax, (int)V, pl.VP, pl.Spad).total; };
  const int kforce = env_int("SSD_LOSS_K", 0);
  int ctas = env_int("SSD_LOSS_CTAS", 0);
  int K = 0;
  if (kforce > 0) {
    K = kforce > 32 ? 32 : kforce;
    if (bytes(K) > cap) return false;
    if (ctas <= 0) ctas = (int)(cap / (bytes(K) + 1024));
  } else {
    // prefer two resident CTAs per SM (8 busy warps) when that still allows K >= 6
    const int want = ctas > 0 ? ctas : 2;
    for (int c = want; c >= 1 && K == 0; --c) {
      int k = 16;
      while (k >= 2 && (bytes(k) + 1024) * c > cap) --k;
      if (k >= (c > 1 ? 6 : 2)) {
        K = k;
        ctas = c;
      }
    }
    if (K == 0) return false;
  }
  if (ctas < 1) ctas = 1;
  pl.K = K;
  pl.ctas_per_sm = ctas;
  pl.smem = bytes(K);
  const int64_t slots = (int64_t)device_sms() * ctas;
  pl.grid = (int)(B < slots ? B : slots);
  pl.nseg_max = (int)((T / 2 + 1 + K - 1) / K + 2);
  pl.ckpt_cta_bytes = (size_t)2 * pl.nseg_max * ((size_t)pl.NCmax * 32 * sizeof(R) + 16);
  pl.ws_bytes = 256 + (size_t)pl.grid * pl.ckpt_cta_bytes;
  return true;
}

static bool loss_use_float() { return env_int("SSD_LOSS_FP32", 0) != 0; }

size_t loss_workspace_bytes(int64_t B, int64_t T, int64_t V, int64_t S_max) {
  LossPlan pl;
  const bool ok = loss_use_float() ? plan_for<float>(B, T, V, S_max, pl) : plan_for<double>(B, T, V, S_max, pl);
  return ok ? pl.ws_bytes : 256;
}

template <typename R>
static int loss_launch_t(LossParams& P, const LossPlan& pl, cudaStream_t stream) {
#ifndef SSD_EMU
  cudaError_t e = cudaFuncSetAttribute(loss_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
  if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "loss: cudaFuncSetAttribute(%zu): %s", pl.smem, cudaGetErrorString(e));
  e = cudaMemsetAsync(P.counter, 0, 256, stream);
  if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "loss: cudaMemsetAsync: %s", cudaGetErrorString(e));
#else
  *P.counter = 0u;
#endif
  SSD_LAUNCH(loss_kernel<R>, (unsigned)pl.grid, LS_THREADS, pl.smem, stream, P);
  return check_launch("loss_kernel");
}

int loss_launch(const float* lp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st, const int64_t* targets,
                int64_t tgt_stride, int64_t S_max, const int64_t* in_len, const int64_t* tg_len, const int32_t* order,
                int blank, int zero_inf, int64_t mean_den, const float* grad_mul, float* nll, float* grad, void* ws,
                size_t ws_bytes, cudaStream_t stream) {
  if (B == 0) return SSD_OK;
  const bool f32 = loss_use_float();
  LossPlan pl;
  const bool ok = f32 ? plan_for<float>(B, T, V, S_max, pl) : plan_for<double>(B, T, V, S_max, pl);
  if (!ok) return fail(SSD_ERR_UNSUPPORTED, "ssd_ctc_loss: no shared-memory plan for V=%lld S_max=%lld", (long long)V, (long long)S_max);
  if (!ws || ws_bytes < pl.ws_bytes) return fail(SSD_ERR_WORKSPACE, "ssd_ctc_loss: workspace %zu bytes, need %zu", ws_bytes, pl.ws_bytes);
  LossParams P;
  P.lp = lp; P.T = T; P.sb = sb; P.st = st; P.B = (int)B; P.V = (int)V;
  P.targets = targets; P.tgt_stride = tgt_stride; P.S_max = (int)S_max;
  P.in_len = in_len; P.tg_len = tg_len; P.order = order;
  P.blank = blank; P.zero_inf = zero_inf; P.mean_den = mean_den; P.grad_mul = grad_mul;
  P.nll = nll; P.grad = grad;
  P.counter = reinterpret_cast<unsigned*>(ws);
  P.ckpt = reinterpret_cast<unsigned char*>(ws) + 256;
  P.ckpt_cta_bytes = pl.ckpt_cta_bytes; P.nseg_max = pl.nseg_max;
  P.K = pl.K; P.NCmax = pl.NCmax; P.VP = pl.VP; P.Spad = pl.Spad;
  P.use_tma = (st == V) && aligned16(lp) && ((sb * 4) % 16 == 0) && ((V * 4) % 16 == 0) ? 1 : 0;
  return f32 ? loss_launch_t<float>(P, pl, stream) : loss_launch_t<double>(P, pl, stream);
}

}  // namespace ssd
