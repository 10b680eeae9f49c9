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

}  // namespace bathgpu
