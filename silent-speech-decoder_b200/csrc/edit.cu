// Levenshtein alignment counts for WER / CER, one warp per (reference, hypothesis) pair.
//
// Replaces the arithmetic of reference src/evaluation/evaluate.py:34-38 (jiwer.wer / jiwer.cer: total
// edits over total reference tokens) and :61-83 (`_levenshtein_counts`: the alignment with the fewest
// edits and, among those, the most hits).  That order is lexicographic on (cost, -hits), which is an
// ordered additive key, so one integer per cell is enough:  key = cost * 2^16 - hits,
//     key[i][j] = min(key[i][j-1] + 2^16, key[i-1][j] + 2^16, key[i-1][j-1] + (r_i == h_j ? -1 : 2^16)).
// Substitutions, deletions and insertions follow from (cost, hits, n, m) alone:
//     n = S + D + H,  m = S + I + H,  cost = S + D + I.
//
// Mapping: a lane owns ED_CW consecutive hypothesis columns in registers; the warp sweeps the rows
// skewed by one step per lane (lane l works on row s - l at step s), so the only exchange is one
// shuffle per step handing a lane's last column to its right neighbour.  Hypotheses longer than
// 32 * ED_CW columns are processed in column tiles; the column on a tile's right edge is kept in
// shared memory (one int per reference row) and is the next tile's left edge.
#include "ssd_common.cuh"

namespace ssd {

constexpr int ED_CW = 8;
constexpr int ED_TILE = 32 * ED_CW;
constexpr int ED_WARPS = 4;
constexpr int ED_K = 1 << 16;

__global__ void __launch_bounds__(ED_WARPS * 32) edit_kernel(const int32_t* __restrict__ ref_tok, const int64_t* __restrict__ ref_off,
                                                            const int32_t* __restrict__ hyp_tok, const int64_t* __restrict__ hyp_off,
                                                            int64_t N, int edge_stride, int32_t* __restrict__ counts) {
  SSD_DYN_SMEM(smem_raw);
  const int lane = (int)threadIdx.x & 31, warp = (int)threadIdx.x >> 5;
  const int64_t pair = (int64_t)blockIdx.x * ED_WARPS + warp;
  if (pair >= N) return;  // whole warps leave; nothing below synchronises across warps
  int* edge = reinterpret_cast<int*>(smem_raw) + (size_t)warp * edge_stride;

  const int32_t* r = ref_tok + ref_off[pair];
  const int32_t* h = hyp_tok + hyp_off[pair];
  const int n = (int)(ref_off[pair + 1] - ref_off[pair]);
  const int m = (int)(hyp_off[pair + 1] - hyp_off[pair]);

  int key = 0;
  if (n == 0 || m == 0) {
    key = (n + m) * ED_K;  // all insertions or all deletions
  } else {
    for (int j0 = 0; j0 < m; j0 += ED_TILE) {
      const int c0 = j0 + lane * ED_CW;  // columns c0+1 .. c0+ED_CW of the table
      int hv[ED_CW], up[ED_CW];
#pragma unroll
      for (int c = 0; c < ED_CW; ++c) {
        hv[c] = c0 + c < m ? h[c0 + c] : -1;  // beyond the hypothesis: never matches (ids are >= 0)
        up[c] = (c0 + c + 1) * ED_K;          // row 0: insertions only
      }
      int corner = c0 * ED_K;  // key[i-1][c0], starts as row 0
      int last = 0;            // my right-most column of the row I finished in the previous step
      const bool first_tile = j0 == 0, more = j0 + ED_TILE < m;
      const int steps = n + 31;
      for (int s = 1; s <= steps; ++s) {
        const int from_left = __shfl_up_sync(FULL, last, 1);
        const int i = s - lane;
        if (i >= 1 && i <= n) {
          const int ri = r[i - 1];
          int left = lane ? from_left : (first_tile ? i * ED_K : edge[i]);
          int diag = corner;
          corner = left;
#pragma unroll
          for (int c = 0; c < ED_CW; ++c) {
            int cur = min(left, up[c]) + ED_K;
            cur = min(cur, diag + (ri == hv[c] ? -1 : ED_K));
            diag = up[c];
            up[c] = cur;
            left = cur;
          }
          last = left;
          if (more && lane == 31) edge[i] = last;  // read again by lane 0 of the next tile, 31 steps after its own read of edge[i]
          if (i == n) {
#pragma unroll
            for (int c = 0; c < ED_CW; ++c)
              if (c0 + c + 1 == m) key = up[c];
          }
        }
      }
      __syncwarp();
    }
    // exactly one lane of one tile saw column m at row n
    key = __shfl_sync(FULL, key, ((m - 1) % ED_TILE) / ED_CW);
  }
  if (lane == 0) {
    const int cost = (key + ED_K - 1) >> 16;
    const int hits = cost * ED_K - key;
    int32_t* o = counts + pair * 4;
    o[0] = (n - hits) + (m - hits) - cost;  // substitutions
    o[1] = cost - (m - hits);               // deletions
    o[2] = cost - (n - hits);               // insertions
    o[3] = hits;
  }
}

int edit_launch(const int32_t* ref_tok, const int64_t* ref_off, const int32_t* hyp_tok, const int64_t* hyp_off, int64_t N,
                int64_t max_ref_len, int64_t max_hyp_len, int32_t* counts, cudaStream_t stream) {
  if (N == 0) return SSD_OK;
  // the edge column is only touched when some hypothesis spans more than one tile
  const int edge_stride = max_hyp_len > ED_TILE ? (int)max_ref_len + 1 : 0;
  const size_t smem = (size_t)ED_WARPS * edge_stride * sizeof(int);
#ifndef SSD_EMU
  {  // per launch: the attribute belongs to the current device, and one process may drive several
    cudaError_t e = cudaFuncSetAttribute(edit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         ED_WARPS * (SSD_MAX_EDIT_LEN + 1) * (int)sizeof(int));
    if (e != cudaSuccess) return fail(SSD_ERR_CUDA, "edit: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
#endif
  const unsigned grid = (unsigned)((N + ED_WARPS - 1) / ED_WARPS);
  SSD_LAUNCH(edit_kernel, grid, ED_WARPS * 32, smem, stream, ref_tok, ref_off, hyp_tok, hyp_off, N, edge_stride, counts);
  return check_launch("edit_kernel");
}

}  // namespace ssd
