// Isolation microbenchmark: which part of the sweep step limits it?  (see ubench_fast.cu for the build line)
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include <cstdint>

constexpr unsigned FULL = 0xffffffffu;
template <int NC, bool LDS, bool SHFL, bool MATH>
__global__ void bench(const float* pglob, int nsteps, long long* out, float* sink) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 49 * 9; i += blockDim.x) sm[i] = pglob[i % 392];
  __syncthreads();
  uint32_t ptr[NC / 2 + 1];
  float skf[NC / 2];
  for (int k = 0; k <= NC / 2; ++k) ptr[k] = (uint32_t)__cvta_generic_to_shared(sm) + ((lane * 7 + k * 3) % 48) * 36;
  for (int k = 0; k < NC / 2; ++k) skf[k] = (k + lane) & 1 ? 1.f : 0.f;
  float x[NC];
  for (int i = 0; i < NC; ++i) x[i] = 1.f + 0.01f * i + lane;
  float fin = 0.9f, inflow = 1.f;
  float p[NC / 2 + 1];
  for (int k = 0; k <= NC / 2; ++k) p[k] = 0.4f + 0.01f * k;
  long long t0 = clock64();
#pragma unroll 1
  for (int t = 0; t < nsteps; ++t) {
    if (LDS) {
#pragma unroll
      for (int k = 0; k <= NC / 2; ++k) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(p[k]) : "r"(ptr[k] + (t & 7) * 4));
    }
    float nxt = inflow;
    if (MATH) {
      {
        float sum = x[NC - 1] + x[NC - 2];
        sum = fmaf(skf[NC / 2 - 1], x[NC - 3], sum);
        x[NC - 1] = sum * p[NC / 2 - 1];
      }
      if (SHFL) nxt = __shfl_up_sync(FULL, x[NC - 1], 1);
#pragma unroll
      for (int i = NC - 2; i >= 2; --i) {
        float sum = x[i] + x[i - 1];
        if (i & 1) { sum = fmaf(skf[i >> 1], x[i - 2], sum); x[i] = sum * p[i >> 1]; }
        else x[i] = sum * p[NC / 2];
      }
      const float n1 = inflow * fin;
      { float sum = x[1] + x[0]; sum = fmaf(skf[0], n1, sum); x[1] = sum * p[0]; }
      x[0] = (x[0] + n1) * p[NC / 2];
    } else {
      if (SHFL) nxt = __shfl_up_sync(FULL, x[NC - 1] + inflow, 1);
      x[0] += p[0] + p[NC / 2];
    }
    inflow = nxt;
  }
  long long t1 = clock64();
  float acc = inflow;
  for (int i = 0; i < NC; ++i) acc += x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (lane == 0) out[blockIdx.x * (blockDim.x / 32) + warp] = t1 - t0;
}

template <int NC, bool LDS, bool SHFL, bool MATH>
void run(const char* name) {
  float* p; long long* out; float* sink;
  cudaMalloc(&p, 392 * 4);
  std::vector<float> h(392);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 0.9f + 0.001f * (i % 37);
  cudaMemcpy(p, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&sink, 1024 * 32 * 4);
  const int ns = 4000;
  for (int warps : {1, 4, 16}) {
    bench<NC, LDS, SHFL, MATH><<<1, 32 * warps, 49 * 9 * 4>>>(p, ns, out, sink);
    bench<NC, LDS, SHFL, MATH><<<1, 32 * warps, 49 * 9 * 4>>>(p, ns, out, sink);
    cudaDeviceSynchronize();
    long long ho[64];
    cudaMemcpy(ho, out, warps * 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps %2d: %7.1f cycles/step  (%s)\n", name, warps, (double)ho[0] / ns, cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  run<14, true, true, true>("NC14 lds+shfl+math");
  run<14, false, true, true>("NC14 shfl+math");
  run<14, true, false, true>("NC14 lds+math");
  run<14, false, false, true>("NC14 math only");
  run<14, true, false, false>("NC14 lds only");
  run<14, false, true, false>("NC14 shfl only");
  return 0;
}
