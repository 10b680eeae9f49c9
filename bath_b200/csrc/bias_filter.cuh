// bias_filter.cuh -- the 2-state bias-composition filter: esl_hmm_Forward over the filter HMM that p7_bg_SetFilter
// builds, as p7_bg_FilterScore runs it on an ORF and p7_bg_fs_FilterScore on the three reading frames of a DNA window
// (src/p7_bg.c:449-471, :491-500, :522-573; Easel esl_hmm.c).  A batch is 10^4-10^5 independent chains of L steps; one warp per chain
// (see the kernel).  Every operation is the host restatement's (Background::hmm_forward, bath_b200/host/pipeline.cpp),
// in its order, with the contraction into FMAs switched off by writing the roundings out (__fmul_rn / __fadd_rn / __fdiv_rn), so the
// scores are bit-identical to the host's and the filter decisions built on them cannot differ.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bathgpu {

struct BiasItem {          // device copy of bathgpu_bias_item
  long long start;         // kind 0: offset of the ORF's first residue in the slot's residue buffer; kind 1: slot coordinate (1-based) of the window's first nucleotide
  int       L;             // residues / nucleotides
  int       table;         // emission-odds table to use
  float     t00;           // p1 = L'/(L'+1) of the null model's length (p7_bg_SetLength copies it into the filter HMM)
  int       pad_;
};

struct BiasArgs {
  const BiasItem *items;
  int             n;
  int             kind;            // 0: ORF residues; 1: DNA window, frames 1..3 (three results per item)
  const float    *tables;          // [ntab][29][2] emission odds e[k][x] / f[x] (esl_hmm_Configure)
  float           t10, t11;        // 1/(L1+1), L1/(L1+1), L1 = M/8
  const uint8_t  *residues;        // kind 0
  const uint32_t *dna4;            // kind 1: 4-bit packed, nt p (0-based) in word (p+8)>>3
  uint8_t         gcode[64];       // kind 1: amino code of codon 16a+4b+c (27 = stop)
  float          *out;             // [n] or [3n]: sum of the log scale factors (what esl_hmm_Forward returns)
};

__device__ __forceinline__ int bias_nt(const uint32_t *__restrict__ dna4, long long p0)      // nucleotide code at 0-based slot position p0
{
  const long long q = p0 + 8;
  return (int)((__ldg(dna4 + (q >> 3)) >> (4 * (int)(q & 7))) & 15u);
}

// One WARP per ORF / per (window, frame).  The recursion itself is a short serial chain per residue (a dozen float operations and two
// divisions); what made the one-thread-per-item version slow is the double-precision log of every row maximum on that chain.  All 32
// lanes run the (cheap, identical) recursion, lane l keeps the maximum of step 32 c + l, the 32 logs of a chunk are taken in parallel,
// and the float sum is then accumulated in step order -- the same operations on the same operands in the same order as the host's
// loop, so the result is still bit-identical to it.
__global__ void __launch_bounds__(128) bias_forward_kernel(BiasArgs a)
{
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int per = (a.kind == 1) ? 3 : 1;
  if (wid >= (long long)a.n * per) return;
  const int it = (int)(wid / per), fr = (int)(wid % per);       // frame fr+1 starts at window position fr+1
  const BiasItem d = a.items[it];
  const float *__restrict__ eo = a.tables + (size_t)d.table * 58;
  const float t00 = d.t00, t01 = __fadd_rn(1.0f, -t00), t10 = a.t10, t11 = a.t11;
  float p0 = 0.f, p1 = 0.f, logsc = 0.f;
  bool  first = true;
  const int nstep = (a.kind == 1) ? (d.L - 2 - fr + 2) / 3 : d.L;     // codons starting at fr+1, fr+4, ... <= L-2
  float mymx = 1.0f;                                              // the row maximum this lane takes the log of
  int   filled = 0;                                               // maxima waiting in the current chunk
  for (int s = 0; s <= nstep; ++s) {
    bool have = false;
    float mx = 1.0f;
    if (s < nstep) {
      int x = -1;
      if (a.kind == 0) x = a.residues[d.start + s];
      else {
        const long long p = d.start - 1 + fr + 3LL * s;          // 0-based slot position of the codon's first nucleotide
        const int n1 = bias_nt(a.dna4, p), n2 = bias_nt(a.dna4, p + 1), n3 = bias_nt(a.dna4, p + 2);
        if (n1 < 4 && n2 < 4 && n3 < 4) { x = a.gcode[16 * n1 + 4 * n2 + n3]; if (x >= 20) x = -1; }   // canonical residues only (p7_bg.c:548-551)
      }
      if (x >= 0) {
        const float e0 = __ldg(eo + 2 * x), e1 = __ldg(eo + 2 * x + 1);
        float c0, c1;
        if (first) { c0 = __fmul_rn(e0, 0.999f); c1 = __fmul_rn(e1, 0.001f); first = false; }
        else {
          c0 = __fmul_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(p0, t00)), __fmul_rn(p1, t10)), e0);
          c1 = __fmul_rn(__fadd_rn(__fadd_rn(0.0f, __fmul_rn(p0, t01)), __fmul_rn(p1, t11)), e1);
        }
        mx = fmaxf(c0, 0.0f); mx = fmaxf(c1, mx);
        p0 = __fdiv_rn(c0, mx); p1 = __fdiv_rn(c1, mx);
        have = true;
      }
    } else if (!first) {                                          // termination row: log(sum of the final state values)
      mx = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(p0, 1.0f)), __fmul_rn(p1, 1.0f));
      have = true;
    }
    if (have) { if (lane == filled) mymx = mx; ++filled; }
    if (filled == 32 || (s == nstep && filled > 0)) {             // warp-uniform: every lane counts the same steps
      const float lg = (float)log((double)mymx);
      for (int l = 0; l < filled; ++l) logsc = __fadd_rn(logsc, __shfl_sync(0xffffffffu, lg, l));
      filled = 0;
    }
  }
  if (lane == 0) a.out[wid] = first ? 0.0f : logsc;               // no residue at all: the host returns 0 without a recursion
}

}  // namespace bathgpu
