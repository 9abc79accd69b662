// Micro-benchmark of the CTC recursion step in isolation: cycles per frame for one warp (no pipeline,
// no barriers), to separate arithmetic/shuffle latency from pipeline effects.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include -o /tmp/ubench profiles/ubench_step.cu
#include <cstdio>
#include "../silent-speech-decoder_b200/csrc/ctc_loss.cu"

namespace ssd { char* error_buffer() { static thread_local char b[512]; return b; } }
using namespace ssd;

template <int NC, int DIR, bool STORE>
__global__ void k_step(const double* prows, int VP, int T, long long* cycles, double* sink, int* wbuf) {
  extern __shared__ unsigned char sm[];
  double* p = reinterpret_cast<double*>(sm);
  for (int i = threadIdx.x; i < 8 * VP; i += blockDim.x) p[i] = prows[i];
  __syncthreads();
  int lam[NC / 2];
  for (int k = 0; k < NC / 2; ++k) lam[k] = ((threadIdx.x * 7 + k * 3) % 46 + 2) * 8;
  PRow<NC> pc;
  load_prow<NC>(pc, sm, lam, 8);
  unsigned skipm = 0x55555555u;
  Strip<NC> s;
  sweep_init<NC>(DIR, 32 * NC - 1, s);
  int* wdst = wbuf + (threadIdx.x >> 5) * NC * 32 + (threadIdx.x & 31);
  long long t0 = clock64();
  for (int t = 0; t < T; ++t) {
    ctc_step<NC, DIR, STORE>(s, skipm, pc, wdst);
    load_prow<NC>(pc, sm + ((t + 1) & 7) * VP * 8, lam, 8);
    if ((t & (LS_RESCALE - 1)) == LS_RESCALE - 1) { renorm<NC>(s); exchange<NC, DIR>(s); }
  }
  long long t1 = clock64();
  double acc = 0;
  for (int i = 0; i < NC; ++i) acc += s.x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + s.e;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NC, int DIR, bool STORE>
void run(const char* name, int warps, int blocks) {
  const int VP = 50, T = 4000;
  double* h = new double[8 * VP];
  for (int i = 0; i < 8 * VP; ++i) h[i] = 0.02 + 0.001 * (i % 17);
  double *d, *sink; long long* cyc; int* wbuf;
  cudaMalloc(&d, 8 * VP * 8); cudaMemcpy(d, h, 8 * VP * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&sink, blocks * warps * 32 * 8); cudaMalloc(&cyc, blocks * 8); cudaMalloc(&wbuf, 64 * 32 * 32 * 4);
  k_step<NC, DIR, STORE><<<blocks, warps * 32, 8 * VP * 8>>>(d, VP, T, cyc, sink, wbuf);
  cudaDeviceSynchronize();
  k_step<NC, DIR, STORE><<<blocks, warps * 32, 8 * VP * 8>>>(d, VP, T, cyc, sink, wbuf);
  cudaError_t e = cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/block %d blocks %3d : %7.1f cycles/frame  (%s)\n", name, warps, blocks, (double)c / T, cudaGetErrorString(e));
  cudaFree(d); cudaFree(sink); cudaFree(cyc); cudaFree(wbuf); delete[] h;
}

int main() {
  run<4, 0, false>("fwd NC=4", 1, 1);
  run<8, 0, false>("fwd NC=8", 1, 1);
  run<14, 0, false>("fwd NC=14", 1, 1);
  run<8, 1, false>("bwd NC=8", 1, 1);
  run<14, 1, false>("bwd NC=14", 1, 1);
  run<8, 0, true>("fwd NC=8 store", 1, 1);
  run<14, 0, true>("fwd NC=14 store", 1, 1);
  run<8, 0, false>("fwd NC=8 4 warps", 4, 1);
  run<8, 0, false>("fwd NC=8 8 warps", 8, 1);
  run<14, 0, false>("fwd NC=14 8 warps", 8, 1);
  run<14, 0, false>("fwd NC=14 8 warps x296", 8, 296);
  return 0;
}
