// Which part of the recursion step is latency-critical?  Self-contained variants.
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
template <typename R, int NC, bool SHFL, bool ALIGN, bool RENORM, bool LDSP>
__global__ void k(const R* prows, int VP, int T, long long* cycles, R* sink) {
  extern __shared__ unsigned char sm[];
  R* p = reinterpret_cast<R*>(sm);
  for (int i = threadIdx.x; i < 8 * VP; i += blockDim.x) p[i] = prows[i];
  __syncthreads();
  int lam[NC / 2];
  for (int k = 0; k < NC / 2; ++k) lam[k] = ((threadIdx.x * 7 + k * 3) % 46 + 2);
  unsigned skipm = 0x55555555u;
  R x[NC];
  for (int i = 0; i < NC; ++i) x[i] = (threadIdx.x == 0 && i == 0) ? R(1) : R(0);
  int e = 0, eseen = -1; R fin = 1; bool z = threadIdx.x != 0;
  const int lane = threadIdx.x & 31;
  long long t0 = clock64();
  for (int t = 0; t < T; ++t) {
    const R* prow = p + (t & 7) * VP;
    R n0 = R(0);
    if (SHFL) { n0 = __shfl_up_sync(FULL, x[NC - 1], 1); if (lane == 0) n0 = 0; }
    if (ALIGN) {
      int ein = __shfl_up_sync(FULL, e, 1);
      if (n0 != R(0)) {
        if (z) { e = ein; z = false; fin = 1; eseen = ein; }
        else { if (ein != eseen) { int d = ein - e; fin = d < -100 ? R(0) : (R)ldexp(1.0, d); eseen = ein; } n0 *= fin; }
      }
    }
    const R pb = LDSP ? prow[1] : R(0.02);
#pragma unroll
    for (int i = NC - 1; i >= 0; --i) {
      const R a1 = (i >= 1) ? x[i >= 1 ? i - 1 : 0] : n0;
      R s = x[i] + a1;
      if (i & 1) {
        const R a2 = (i >= 2) ? x[i >= 2 ? i - 2 : 0] : n0;
        if ((skipm >> (i >> 1)) & 1u) s += a2;
        x[i] = s * (LDSP ? prow[lam[i >> 1]] : R(0.021));
      } else x[i] = s * pb;
    }
    if (RENORM && (t & 3) == 3) {
      R m = 0;
#pragma unroll
      for (int i = 0; i < NC; ++i) m = x[i] > m ? x[i] : m;
      int ex; frexp((double)m, &ex);
      const R f = m == R(0) ? R(1) : (R)ldexp(1.0, -ex);
#pragma unroll
      for (int i = 0; i < NC; ++i) x[i] *= f;
      e += ex; eseen = -1;
    } else if (!RENORM && (t & 3) == 3) {
#pragma unroll
      for (int i = 0; i < NC; ++i) x[i] *= R(1e4);   // keep values alive without the reduction
    }
  }
  long long t1 = clock64();
  R acc = 0;
  for (int i = 0; i < NC; ++i) acc += x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + e;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <typename R, int NC, bool SHFL, bool ALIGN, bool RENORM, bool LDSP>
void run(const char* name) {
  const int VP = 50, T = 4000;
  R* h = new R[8 * VP];
  for (int i = 0; i < 8 * VP; ++i) h[i] = R(0.02 + 0.001 * (i % 17));
  R *d, *sink; long long* cyc;
  cudaMalloc(&d, 8 * VP * sizeof(R)); cudaMemcpy(d, h, 8 * VP * sizeof(R), cudaMemcpyHostToDevice);
  cudaMalloc(&sink, 32 * sizeof(R)); cudaMalloc(&cyc, 8);
  for (int rep = 0; rep < 2; ++rep) { k<R, NC, SHFL, ALIGN, RENORM, LDSP><<<1, 32, 8 * VP * sizeof(R)>>>(d, VP, T, cyc, sink); cudaDeviceSynchronize(); }
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-46s %7.1f cycles/frame\n", name, (double)c / T);
}
int main() {
  run<double, 8, true, true, true, true>("f64 NC8 shfl+align+renorm+lds");
  run<double, 8, true, true, false, true>("f64 NC8 shfl+align       +lds");
  run<double, 8, true, false, false, true>("f64 NC8 shfl             +lds");
  run<double, 8, false, false, false, true>("f64 NC8                  +lds");
  run<double, 8, false, false, false, false>("f64 NC8 arithmetic only");
  run<double, 8, true, false, false, false>("f64 NC8 shfl only");
  run<double, 8, false, false, true, false>("f64 NC8 renorm only");
  run<float, 8, true, true, true, true>("f32 NC8 shfl+align+renorm+lds");
  run<float, 8, false, false, false, false>("f32 NC8 arithmetic only");
  run<double, 14, false, false, false, false>("f64 NC14 arithmetic only");
  run<double, 4, false, false, false, false>("f64 NC4 arithmetic only");
  return 0;
}
