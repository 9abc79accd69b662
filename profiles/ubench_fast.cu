// Microbenchmark of the fp32 sweep step of ctc_loss_fast.cu in isolation (no pipeline around it):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I silent-speech-decoder_b200/csrc -I include \
//        -o profiles/ubench_fast profiles/ubench_fast.cu && profiles/ubench_fast
// Prints cycles per frame for one warp alone on an SM sub-partition and with 1..3 co-resident copies.
#include <cstdio>
#include <vector>
#include "../silent-speech-decoder_b200/csrc/ctc_loss_fast.cu"

namespace ssd { char* error_buffer() { static char b[512]; return b; } }
using namespace ssd;

template <int NC, int DIR, int MODE>
__global__ void bench(const float* pglob, int nblocks, long long* out, float* sink) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int V = 48;
  for (int i = threadIdx.x; i < (V + 1) * F_KS; i += blockDim.x) sm[i] = i < V * F_KS ? pglob[i % 392] : 0.f;
  __syncthreads();
  FLabels<NC> lt;
  for (int k = 0; k < NC / 2; ++k) {
    lt.off[k] = ((lane * 7 + k * 3) % V) * F_KS * 4;
    lt.skf[k] = (k + lane) & 1 ? 1.f : 0.f;
  }
  FStrip<NC> s;
  sweep_init_f<NC, DIR>(2 * 200 + 1, s);
  if (DIR == 0) s.x[0] = 1.f;
  for (int i = 0; i < NC; ++i) s.x[i] = 1.f + 0.01f * i;
  Inflow in;
  inflow_prime<NC, DIR>(s, in);
  const int blank_off = 1 * F_KS * 4;
  long long t0 = clock64();
  for (int b = 0; b < nblocks; ++b) {
    for_frames<NC, DIR>(sm, 0, F_K, lt, blank_off, [&](int, const FProbs<NC>& pc, auto&& mid) { fstep<NC, DIR, false>(s, in, lt, pc, nullptr, mid); });
    if (MODE >= 1) {
      renorm_exchange<NC, DIR>(s);
      inflow_prime<NC, DIR>(s, in);
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
  for (int i = 0; i < NC; ++i) acc += s.x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (lane == 0) out[blockIdx.x * (blockDim.x / 32) + warp] = t1 - t0;
}

template <int NC, int DIR, int MODE>
void run(const char* name, int warps) {
  float* p;
  long long* out;
  float* sink;
  cudaMalloc(&p, 49 * 8 * 4);
  std::vector<float> h(49 * 8);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 0.3f + 0.001f * (i % 37);
  cudaMemcpy(p, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 1024 * 8);
  cudaMalloc(&sink, 1024 * 32 * 4);
  const int nb = 200;
  bench<NC, DIR, MODE><<<1, 32 * warps, 49 * 9 * 4>>>(p, nb, out, sink);
  bench<NC, DIR, MODE><<<1, 32 * warps, 49 * 9 * 4>>>(p, nb, out, sink);
  cudaDeviceSynchronize();
  long long ho[64];
  cudaMemcpy(ho, out, warps * 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps %2d (per SMSP %d): %6.1f cycles/frame (warp 0)  err=%s\n", name, warps, (warps + 3) / 4, (double)ho[0] / (nb * F_K),
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(p); cudaFree(out); cudaFree(sink);
}

int main() {
  for (int w : {1, 4, 8, 16}) run<14, 0, 0>("fwd NC=14 steps only", w);
  for (int w : {1, 4, 8}) run<14, 1, 0>("bwd NC=14 steps only", w);
  for (int w : {1, 4}) run<14, 0, 1>("fwd NC=14 + renorm/8", w);
  for (int w : {1, 4}) run<8, 0, 0>("fwd NC=8 steps only", w);
  return 0;
}
