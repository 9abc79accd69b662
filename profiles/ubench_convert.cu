// Microbenchmark: cycles per convert_block call (8 frames x 48 classes, shared -> class-major shared) for one warp,
// and of the mbarrier ping-pong between two warps (arrive -> try_wait wake-up), the two per-block costs of a loader.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I silent-speech-decoder_b200/csrc -I include \
//        -o profiles/ubench_convert profiles/ubench_convert.cu && profiles/ubench_convert
#include <cstdio>
#include "../silent-speech-decoder_b200/csrc/ctc_loss_fast.cu"
namespace ssd { char* error_buffer() { static char b[512]; return b; } }
using namespace ssd;

__global__ void conv(const float* g, long long* out, int reps, int mode) {
  __shared__ __align__(16) float raw[F_K * 48];
  __shared__ __align__(16) float pblk[49 * F_KS + 8];
  for (int i = threadIdx.x; i < F_K * 48; i += 32) raw[i] = g[i];
  __syncwarp();
  int bad = 0;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    bad |= mode == 0 ? convert_block(raw, g, 48, F_K, 48, pblk) : convert_block_t<true>(raw, g, 48, F_K, 48, pblk);
    __syncwarp();
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / reps, out[1] = bad;
}

__global__ void pingpong(long long* out, int reps) {
  __shared__ __align__(8) uint64_t a, b;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&a, 1); mbar_init(&b, 1); mbar_fence_init(); }
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (warp == 0) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&a);
      mbar_wait(&b, r & 1);
    } else {
      mbar_wait(&a, r & 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&b);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (t1 - t0) / reps;
}

int main() {
  float* g; long long* out; long long h[2];
  cudaMalloc(&g, 8 * 48 * 4); cudaMalloc(&out, 16);
  float hg[8 * 48]; for (int i = 0; i < 8 * 48; ++i) hg[i] = -3.f - 0.01f * (i % 50);
  cudaMemcpy(g, hg, sizeof hg, cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 2; ++mode) {
    conv<<<1, 32>>>(g, out, 200, mode); conv<<<1, 32>>>(g, out, 200, mode);
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("convert_block (%s): %lld cycles per block  [%s]\n", mode ? "inlined" : "noinline", h[0], cudaGetErrorString(cudaGetLastError()));
  }
  pingpong<<<1, 64>>>(out, 1000); pingpong<<<1, 64>>>(out, 1000);
  cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
  printf("mbarrier ping-pong round trip: %lld cycles  [%s]\n", h[0], cudaGetErrorString(cudaGetLastError()));
  return 0;
}
