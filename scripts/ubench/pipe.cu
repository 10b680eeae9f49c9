// scratch micro-benchmark: do warp shuffles and L1-hit loads share one per-SM pipe?  16 warps on one SM, per iteration and warp:
// NS independent SHFLs and NL independent LDG.64 (L1 hits).  If the pipes were separate the time would be max(NS, 2 NL); shared: NS + 2 NL.
#include <cstdio>
#include <cuda_runtime.h>
template <int NS, int NL>
__global__ void k(const float2 *__restrict__ tab, float *out, int iters, long long *cyc)
{
  float v[8]; float2 acc[8];
  for (int u = 0; u < 8; ++u) { v[u] = threadIdx.x + u; acc[u] = make_float2(0.f, 0.f); }
  const float2 *p = tab + (threadIdx.x & 31);
  unsigned off = (threadIdx.x >> 5) * 64;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < NS; ++u) v[u] = __shfl_xor_sync(0xffffffffu, v[u], 1 + (u & 3));
#pragma unroll
    for (int u = 0; u < NL; ++u) { float2 x = __ldg(p + ((off + u * 32 + i * 32) & 1023)); acc[u].x += x.x; acc[u].y += x.y; }
  }
  long long t1 = clock64();
  float s = 0; for (int u = 0; u < 8; ++u) s += v[u] + acc[u].x + acc[u].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int NS, int NL> void run(const float2 *tab, float *out, long long *cyc)
{
  const int iters = 4000, warps = 16; long long c;
  k<NS, NL><<<1, 32 * warps>>>(tab, out, iters, cyc); k<NS, NL><<<1, 32 * warps>>>(tab, out, iters, cyc);
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%d SHFL + %d LDG.64 per warp-iteration, 16 warps: %.2f SM cycles per warp-iteration (separate pipes: %d, one pipe: %d)\n",
         NS, NL, (double)c / (iters * (double)warps), NS > 2 * NL ? NS : 2 * NL, NS + 2 * NL);
}
int main()
{
  float2 *tab; float *out; long long *cyc;
  cudaMalloc(&tab, 2048 * 8); cudaMemset(tab, 0, 2048 * 8); cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  run<8, 0>(tab, out, cyc); run<0, 4>(tab, out, cyc); run<0, 8>(tab, out, cyc); run<8, 4>(tab, out, cyc); run<4, 4>(tab, out, cyc); run<8, 8>(tab, out, cyc); run<4, 8>(tab, out, cyc);
  return 0;
}
