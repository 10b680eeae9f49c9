// microbench.cuh -- FP32 FMA throughput probe: the denominator of the frameshift kernels' roofline.
// MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; the Forward/Backward recursions are bound by the
// FP32 pipe, so the peak they are divided by is measured here, on the same device, same clocks.
#pragma once
#include <cuda_runtime.h>

namespace bathgpu {

// 16 independent FFMA chains per thread (enough ILP to hide the 4-cycle pipe latency with 8+ warps per scheduler)
static __global__ void __launch_bounds__(256) fp32_fma_probe_kernel(float *out, int iters, float a, float b)
{
  float x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = (float)(threadIdx.x + j) * 1e-3f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = fmaf(x[j], a, b);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) s += x[j];
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true; keeps the chains live
}

// 16-bit SIMD-in-register issue probe: the denominator of the integer filters' roofline.  Each statement is one add-then-clamp
// (VIADDMNMX.S16x2) and one max (VIMNMX.S16x2) on both 16-bit halves of a register -- the native instructions the MSV / Viterbi cells
// are made of (orf_filters.cuh); an instruction counts as one operation per half.  16 independent chains per thread.
static __global__ void __launch_bounds__(256) int16x2_probe_kernel(unsigned *out, int iters, unsigned a, unsigned b)
{
  unsigned x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = (threadIdx.x + j) * 0x00010001u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int j = 0; j < 16; ++j) x[j] = __vmaxs2(__viaddmin_s16x2(x[j], a, 0x00ff00ffu), b);
    }
  }
  unsigned s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s ^= x[j];
  if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // never true in practice; keeps the chains live
}

}  // namespace bathgpu
