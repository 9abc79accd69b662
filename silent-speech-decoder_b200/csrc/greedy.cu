// Greedy CTC decode: per-frame argmax + collapse, one CTA per utterance.
//
// Replaces reference src/decoding/ctc.py:16-36 (_greedy_decode).  HBM-bound: every log-prob is
// read exactly once (V*4 bytes per frame) and at most one int32 per frame is written.
//
// Data movement: the (FR x V) fp32 tile of an utterance is contiguous in HBM when
// stride_t == V, so it is staged into shared memory by ONE 1-D TMA bulk copy per tile
// (cp.async.bulk, SASS UBLKCP) signalled on an mbarrier, double buffered so tile k+1 is in
// flight while tile k is reduced.  Each thread then owns one frame and walks its row starting at
// a per-lane rotated column so the 32 lanes of a warp hit different banks.  Non-contiguous or
// unaligned inputs take a cooperative coalesced-load path into the same tiles.
//
// Collapse: frame t emits iff a_t != blank and a_t != a_{t-1} (a blank resets "previous",
// which is the same predicate).  Emit flags are compacted with a ballot scan.
#include "ssd_common.cuh"

namespace ssd {

constexpr int GR_THREADS = 128;  // frames per tile == threads per CTA

__global__ void __launch_bounds__(GR_THREADS) greedy_kernel(const float* __restrict__ lp, int64_t T, int V, int64_t sb, int64_t st,
                                                            const int64_t* __restrict__ lengths, int blank, float bias,
                                                            int32_t* __restrict__ out_tok, int32_t* __restrict__ out_len,
                                                            int use_tma) {
  SSD_DYN_SMEM(smem_raw);
  float* tile0 = reinterpret_cast<float*>(smem_raw);
  const int tile_elems = GR_THREADS * V;
  float* tiles[2] = {tile0, tile0 + tile_elems};
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int warp_cnt[GR_THREADS / 32];
  __shared__ int carry_prev;  // argmax of the last frame of the previous tile

  const int tid = (int)threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int64_t b = (int64_t)blockIdx.x;
  int64_t len64 = lengths[b];
  const int len = (int)(len64 < 0 ? 0 : (len64 > T ? T : len64));
  const float* base = lp + b * sb;
  int32_t* otok = out_tok + b * T;
  const int ntiles = (len + GR_THREADS - 1) / GR_THREADS;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
    carry_prev = -1;
  }
  __syncthreads();

  auto issue = [&](int k) {  // thread 0 only
    const int rows = min(GR_THREADS, len - k * GR_THREADS);
    const uint32_t bytes = (uint32_t)(rows * V * 4);
    mbar_arrive_expect_tx(&bar[k & 1], bytes);
    bulk_g2s(tiles[k & 1], base + (int64_t)k * GR_THREADS * st, bytes, &bar[k & 1]);
  };
  if (use_tma && tid == 0 && ntiles > 0) issue(0);

  int emitted = 0;  // running output count (same in every thread)
  for (int k = 0; k < ntiles; ++k) {
    const int t0 = k * GR_THREADS;
    const int rows = min(GR_THREADS, len - t0);
    float* tile = tiles[k & 1];
    if (use_tma) {
      if (tid == 0 && k + 1 < ntiles) issue(k + 1);  // buffer (k+1)&1 was released by the barrier ending tile k-1
      mbar_wait(&bar[k & 1], (uint32_t)((k >> 1) & 1));
    } else {
      for (int i = tid; i < rows * V; i += GR_THREADS) {
        const int r = i / V, v = i - r * V;
        tile[i] = ld_stream(base + (int64_t)(t0 + r) * st + v);
      }
      __syncthreads();
    }

    // ---- argmax of my frame, first maximum wins.  Two passes over my row in shared memory: the
    // maximum (one FMNMX per element), then the lowest index attaining it.  Rows are V*4 bytes apart,
    // so each lane starts at a different 16-byte chunk to spread the banks.
    int best_i = -1;
    if (tid < rows) {
      const float* row = tile + tid * V;
      // the blank bias is applied to the VALUE as it is compared (ctc.py:18-20 adds it to a clone): the tile is
      // never written by the generic proxy, so the next bulk copy into this buffer needs no cross-proxy fence
      const bool biased = bias != 0.f && blank >= 0 && blank < V;
      if ((V & 3) == 0) {
        const float4* row4 = reinterpret_cast<const float4*>(row);
        const int V4 = V >> 2;
        const int cb = biased ? (blank >> 2) : -1, jb = blank & 3;
        auto fetch = [&](int c) {
          float4 x = row4[c];
          if (c == cb) {
            if (jb == 0) x.x += bias;
            else if (jb == 1) x.y += bias;
            else if (jb == 2) x.z += bias;
            else x.w += bias;
          }
          return x;
        };
        int c = tid % V4;
        float m = -INFINITY;
        for (int j = 0; j < V4; ++j) {
          const float4 x = fetch(c);
          m = fmaxf(fmaxf(m, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
          c = (c + 1 == V4) ? 0 : c + 1;
        }
        int bi = 0x7fffffff;
        for (int j = 0; j < V4; ++j) {
          const float4 x = fetch(c);
          const int i0 = 4 * c;
          // highest index first so that the lowest matching one is what remains
          int k = x.w == m ? i0 + 3 : 0x7fffffff;
          k = x.z == m ? i0 + 2 : k;
          k = x.y == m ? i0 + 1 : k;
          k = x.x == m ? i0 : k;
          bi = k < bi ? k : bi;
          c = (c + 1 == V4) ? 0 : c + 1;
        }
        best_i = bi == 0x7fffffff ? 0 : bi;  // all-NaN row: index 0
      } else {
        float best = 0.f;
        for (int v = 0; v < V; ++v) {
          const float x = row[v] + ((biased && v == blank) ? bias : 0.f);
          if (best_i < 0 || x > best) {
            best = x;
            best_i = v;
          }
        }
      }
    }
    // ---- previous frame's argmax
    int prev = __shfl_up_sync(FULL, best_i, 1);
    __shared__ int warp_last[GR_THREADS / 32];
    if (lane == 31) warp_last[warp] = best_i;
    __syncthreads();  // also: everyone is done reading carry_prev's producer and warp_last is visible
    if (lane == 0) prev = (warp == 0) ? carry_prev : warp_last[warp - 1];
    const bool emit = (tid < rows) && best_i != blank && best_i != prev;
    // ---- compaction
    const unsigned m = __ballot_sync(FULL, emit);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int off = emitted;
    int total = 0;
#pragma unroll
    for (int w = 0; w < GR_THREADS / 32; ++w) {
      const int c = warp_cnt[w];
      if (w < warp) off += c;
      total += c;
    }
    if (emit) otok[off + __popc(m & ((1u << lane) - 1u))] = best_i;
    emitted += total;
    if (tid == rows - 1) carry_prev = best_i;
    __syncthreads();  // tile fully consumed: its buffer may be refilled; carry_prev / warp_cnt settled
  }
  if (tid == 0) out_len[b] = emitted;
}

int greedy_launch(const float* lp, int64_t B, int64_t T, int64_t V, int64_t sb, int64_t st, const int64_t* lengths, int blank,
                  float bias, int32_t* out_tok, int32_t* out_len, cudaStream_t stream) {
  if (B == 0) return SSD_OK;
  const size_t smem = (size_t)2 * GR_THREADS * V * sizeof(float);
  // bulk copies need contiguous rows and 16-byte aligned tile starts and sizes
  const int use_tma = (st == V) && aligned16(lp) && ((sb * 4) % 16 == 0) && ((V * 4) % 16 == 0) ? 1 : 0;
#ifndef SSD_EMU
  {  // per launch: the attribute belongs to the current device, and one process may drive several
    cudaError_t e = cudaFuncSetAttribute(greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * GR_THREADS * SSD_MAX_VOCAB * 4);
    if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "greedy: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
#endif
  SSD_LAUNCH(greedy_kernel, (unsigned)B, GR_THREADS, smem, stream, lp, T, (int)V, sb, st, lengths, blank, bias, out_tok, out_len,
             use_tma);
  return check_launch("greedy_kernel");
}

}  // namespace ssd
