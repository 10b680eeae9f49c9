// scratch micro-benchmark: dependent-chain latencies on sm_100a (SHFL, FFMA, FADD after SHFL, LDG L1 hit), one warp per SM
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_shfl(float *out, int iters, long long *cyc)
{
  float v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) v = __shfl_xor_sync(0xffffffffu, v, 1);
  }
  long long t1 = clock64();
  out[threadIdx.x] = v; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_add(float *out, int iters, long long *cyc)
{
  float v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) v += __shfl_xor_sync(0xffffffffu, v, 1 << (u % 5));
  }
  long long t1 = clock64();
  out[threadIdx.x] = v; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_fma(float *out, int iters, long long *cyc, float a, float b)
{
  float v = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) v = fmaf(v, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = v; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_ldg(const int *p, int *out, int iters, long long *cyc)
{
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) idx = __ldg(p + idx);
  }
  long long t1 = clock64();
  out[threadIdx.x] = idx; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// independent shuffles from one warp: issue rate
__global__ void k_shfl_tput(float *out, int iters, long long *cyc)
{
  float v[8];
  for (int u = 0; u < 8; ++u) v[u] = threadIdx.x + u;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __shfl_xor_sync(0xffffffffu, v[u], 1);
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __shfl_xor_sync(0xffffffffu, v[u], 2);
  }
  long long t1 = clock64();
  float s = 0; for (int u = 0; u < 8; ++u) s += v[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s; if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main()
{
  float *out; int *iout; long long *cyc; int *tab;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&iout, 1 << 20); cudaMalloc(&cyc, 64); cudaMalloc(&tab, 4096 * 4);
  int h[4096]; for (int i = 0; i < 4096; ++i) h[i] = (i * 33 + 7) & 4095;
  cudaMemcpy(tab, h, sizeof h, cudaMemcpyHostToDevice);
  const int iters = 2000; long long c;
  k_shfl<<<1, 32>>>(out, iters, cyc); k_shfl<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent SHFL:            %.2f cycles each\n", (double)c / (iters * 16));
  k_shfl_add<<<1, 32>>>(out, iters, cyc); k_shfl_add<<<1, 32>>>(out, iters, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent SHFL + FADD:     %.2f cycles per step\n", (double)c / (iters * 16));
  k_fma<<<1, 32>>>(out, iters, cyc, 0.999f, 0.001f); k_fma<<<1, 32>>>(out, iters, cyc, 0.999f, 0.001f); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent FFMA:            %.2f cycles each\n", (double)c / (iters * 16));
  k_ldg<<<1, 32>>>(tab, iout, iters, cyc); k_ldg<<<1, 32>>>(tab, iout, iters, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("dependent LDG (L1 hit):    %.2f cycles each\n", (double)c / (iters * 16));
  for (int warps = 1; warps <= 16; warps *= 2) {
    k_shfl_tput<<<1, 32 * warps>>>(out, iters, cyc); k_shfl_tput<<<1, 32 * warps>>>(out, iters, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("independent SHFL, %2d warps on one SM: %.2f cycles per warp-SHFL per SM\n", warps, (double)c / (iters * 16.0 * warps));
  }
  return 0;
}
