// fs_backward.cuh -- frameshift Backward PARSER (3 codon lengths) and domain decoding for sm_100a.
//
// Computes p7_BackwardParser_Frameshift_3Codons (reference: src/impl_sse/fwdback_fs.c:565-1013) and
// p7_DomainDecoding_Frameshift (src/impl_sse/decoding_fs.c:245-359).  Same work decomposition as the
// Forward parser (fs_parser.cuh): one warp per DNA window, lane l owns J contiguous model nodes, all
// per-node state in registers, rows walked from L down to 0 in an unroll of 4.
//
// Reference recurrence for row i (descending), with v(k) = sum_c R[c][k] M(i+c,k), c = 2,3,4 the length of
// the quasi-codon starting at nucleotide i+1:
//     B(i)   = sum_k v(k) tBM(k-1)
//     D(i,k) = E(i) + v(k+1) tDM(k) + D(i,k+1) tDD(k)
//     M(i,k) = E(i) + I(i+3,k) tMI(k) + v(k+1) tMM(k) + D(i,k+1) tMD(k)
//     I(i,k) =        I(i+3,k) tII(k) + v(k+1) tIM(k)
// What the kernel carries instead:
//   * Mt(i,k) = M(i,k) / Z(k) with the same profile constant Z(k) = 1 + tMD(k)(1 + tDD(k+1) + ...) the
//     Forward kernel uses: the E(i) terms of M and of the whole D chain collapse into "E(i) +", so the
//     emission table (already multiplied by tBM(k-1) Z(k)) is shared with Forward, B(i) is a plain sum, and
//     nothing that depends on E(i) sits inside the D chain:  Mt(i,k) = E(i) + G(i,k), where G needs only
//     v, I and the E-free chain D0(i,k) = v(k+1) tDM(k) + D0(i,k+1) tDD(k)  (a downward warp scan);
//   * every chain takes the emission product with coefficient 1, as in the Forward parser (fs_parser.cuh, FwdConsts): this kernel
//     reads its own table copy with vmm(k-1) = tMM(k-1) / (Z(k-1) s(k)) folded into column k, so that the shifted product
//     vn(k) = v(k+1) already is the match->match term of G(k); the D0 chain and the insert row are carried divided by
//     r(k) = vdm(k) / vmm(k) and u(k) = vim(k) / vmm(k):
//         B(i) = sum_k v(k) qb(k)                         qb(k) = 1 / vmm(k-1)        (an FMA where the plain sum needs an add)
//         Ds(k) = vn(k) + Ds(k+1) dd3(k)                  dd3 = r(k+1) tDD(k) / r(k)
//         Is(i,k) = Is(i+3,k) tII(k) + vn(k)
//         G(k) = vn(k) + Is(i+3,k) mi3(k) + Ds(k+1) md3(k)   mi3 = u(k) tMI(k) / Z(k),  md3 = r(k+1) tMD(k) / Z(k)
//     10 floating-point instructions per cell instead of 13 and 5 constants per node instead of 7 (Bck3Consts; the 5-codon
//     Backward of fs_domain.cuh keeps BckConsts);
//   * 4-slot rings for Mt and I (rows i+1..i+4 live), 4-slot rings for the N/J/C specials.
// Scaling follows the reference: row i is divided by the Forward row's SCALE(i) unless Backward has
// switched to its own scales (xB > 1e16; :912-915), the two initialisation rows do not rescale the special
// buffers (:673-678), row 0 is never rescaled.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "fs_parser.cuh"

// The B sums of a row pair in one butterfly (pair_allsum, fs_parser.cuh: 6 shuffles for 10).  Measured on B200 (scripts/j_sweep.py,
// Forward with X rows + Backward, both sweeps counted): +1-2 % at every node count from 2 to 8 (M = 192: 862 -> 880 GCUPS, M = 134: 683 -> 697).
#ifndef BATHGPU_BCK_JOINT_B
#define BATHGPU_BCK_JOINT_B(J) ((J) <= 8)
#endif

namespace bathgpu {

enum BckCellConst { BC_VMM = 0, BC_VIM, BC_VDM, BC_DD, BC_MD, BC_MI, BC_II, BC_COUNT };
enum BckLaneConst { BL_B0 = 0, BL_B1, BL_B2, BL_B3, BL_B4, BL_COUNT };

struct FsBackwardArgs {
  const float    *emis;        // the Backward parser's table copy: R[c][k] tBM(k-1) Z(k) vmm(k-1), permuted (FsProfileImage::emis_bck)
  const float    *cellb;       // Bck3Consts: [B3_COUNT][J][32] + [5][32]
  const float    *cellbmw;     // the same for the multi-warp kernel (fs_backward_mw.cuh; J >= 16), or null
  const uint32_t *dna4;
  const WindowDesc *wins;
  int             nwin;
  int             mpad;
  float           tEM, tEL;
  const float    *fxmx;        // Forward X rows, window w at fxmx + xoff[w]*6
  float          *bxmx;        // Backward X rows (out), same offsets
  const long long *xoff;
  float          *bcksc;       // [nwin]
  int            *status;      // [nwin]  (only written when Backward fails: keeps a Forward failure)
  int            *counter;
};

template <int J>
struct BckConsts {
  float vmm[J], vim[J], vdm[J], dd[J], md[J], mi[J], ii[J];
  float bs[5];
};

enum Bck3CellConst { B3_QB = 0, B3_DD, B3_MD, B3_MI, B3_II, B3_COUNT };

template <int J>
struct Bck3Consts {
  float qb[J], dd[J], md[J], mi[J], ii[J];
  float bs[5];
};

template <int J>
__device__ __forceinline__ void load_bck3_consts(const float *__restrict__ cc, int lane, Bck3Consts<J> &K)
{
#pragma unroll
  for (int j = 0; j < J; ++j) {
    K.qb[j] = __ldg(cc + (B3_QB * J + j) * kWarp + lane);
    K.dd[j] = __ldg(cc + (B3_DD * J + j) * kWarp + lane);
    K.md[j] = __ldg(cc + (B3_MD * J + j) * kWarp + lane);
    K.mi[j] = __ldg(cc + (B3_MI * J + j) * kWarp + lane);
    K.ii[j] = __ldg(cc + (B3_II * J + j) * kWarp + lane);
  }
#pragma unroll
  for (int s = 0; s < 5; ++s) K.bs[s] = __ldg(cc + B3_COUNT * J * kWarp + s * kWarp + lane);
}

template <int J>
struct BckState {
  float Mt[4][J];
  float I[4][J];
  float xN[4], xJ[4], xC[4];
};

template <int J>
__device__ __forceinline__ void load_bck_consts(const float *__restrict__ cc, int lane, BckConsts<J> &K)
{
#pragma unroll
  for (int j = 0; j < J; ++j) {
    K.vmm[j] = __ldg(cc + (BC_VMM * J + j) * kWarp + lane);
    K.vim[j] = __ldg(cc + (BC_VIM * J + j) * kWarp + lane);
    K.vdm[j] = __ldg(cc + (BC_VDM * J + j) * kWarp + lane);
    K.dd[j]  = __ldg(cc + (BC_DD  * J + j) * kWarp + lane);
    K.md[j]  = __ldg(cc + (BC_MD  * J + j) * kWarp + lane);
    K.mi[j]  = __ldg(cc + (BC_MI  * J + j) * kWarp + lane);
    K.ii[j]  = __ldg(cc + (BC_II  * J + j) * kWarp + lane);
  }
#pragma unroll
  for (int s = 0; s < 5; ++s) K.bs[s] = __ldg(cc + BC_COUNT * J * kWarp + (BL_B0 + s) * kWarp + lane);
}

// emission-row indices for Backward row i: quasi-codons STARTING at nucleotide i+1 (fwdback_fs.c:806-818)
__device__ __forceinline__ uint32_t codon_word_bck(uint32_t lo, uint32_t hi, int sh, int i, int L)
{
  uint32_t bits = __funnelshift_r(lo, hi, sh) & 0xffffu;       // nibbles n[i+1], n[i+2], n[i+3], n[i+4]
  int a = (int)(bits & 15u), b = (int)((bits >> 4) & 15u), c = (int)((bits >> 8) & 15u), d = (int)(bits >> 12);
  a = (a < 4 && i + 1 >= 1 && i + 1 <= L) ? a : 338;
  b = (b < 4 && i + 2 >= 1 && i + 2 <= L) ? b : 338;
  c = (c < 4 && i + 3 >= 1 && i + 3 <= L) ? c : 338;
  d = (d < 4 && i + 4 >= 1 && i + 4 <= L) ? d : 338;
  return (uint32_t)codon2_fs3(a, b) | ((uint32_t)codon3_fs3(a, b, c) << 9) | ((uint32_t)codon4_fs3(a, b, c, d) << 18);
}

struct BckRowCtx {
  int   L;
  float ploop, pmove, tEL, tEM;
  float totscale;
  bool  own_scales;
};

// The row itself: emission products, B(i), the scaled D0 chain, G, the insert row, the specials -- everything up to, but not
// including, the rescaling.  Writes Is(i,.) and Mt(i,.) = E(i) + G into slot S0; returns the row's specials, unscaled.
struct BckOut { float xB, xN, xJ, xC, xE; };

template <int J> struct BckMid { float g[J]; float bsum; };

// everything of the row that does not need B(i): emission products, the lane's share of B(i), the scaled D0 chain, G, the insert row
template <int J, int VEC, int PH>
__device__ __forceinline__ void bck_row_front(int lane, BckState<J> &S, const Bck3Consts<J> &K,
                                              const char *__restrict__ emis_lane, unsigned rowbytes, uint32_t cw, BckMid<J> &X)
{
  constexpr int S0 = PH, S2 = (PH + 2) & 3, S3 = (PH + 3) & 3;
  float e2[J], e3[J], e4[J], v[J];
  load_emission_row<J, VEC>(reinterpret_cast<const float *>(emis_lane + (size_t)(cw & 511u) * rowbytes), e2);
  load_emission_row<J, VEC>(reinterpret_cast<const float *>(emis_lane + (size_t)((cw >> 9) & 511u) * rowbytes), e3);
  load_emission_row<J, VEC>(reinterpret_cast<const float *>(emis_lane + (size_t)(cw >> 18) * rowbytes), e4);

  // v(k) = tBM(k-1) vmm(k-1) sum_c R[c][k] M(i+c,k);  B(i) = sum_k v(k) qb(k)            (:820-835)
  float bs0 = 0.f, bs1 = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    float t = S.Mt[S2][j] * e2[j];
    t = fmaf(S.Mt[S3][j], e3[j], t);
    t = fmaf(S.Mt[S0][j], e4[j], t);
    v[j] = t;
    if (j == 0) bs0 = t * K.qb[0]; else if (j == 1) bs1 = t * K.qb[1]; else if (j & 1) bs1 = fmaf(t, K.qb[j], bs1); else bs0 = fmaf(t, K.qb[j], bs0);
  }
  X.bsum = (J > 1) ? bs0 + bs1 : bs0;

  // vn(k) = v(k+1): shift down by one node
  float vn[J];
  {
    float up = __shfl_down_sync(0xffffffffu, v[0], 1);
    if (lane == 31) up = 0.f;
#pragma unroll
    for (int j = 0; j + 1 < J; ++j) vn[j] = v[j + 1];
    vn[J - 1] = up;
  }

  // E-free D chain, downward, scaled: Ds(k) = vn(k) + Ds(k+1) dd3(k)          (:885-909)
  float A = 0.f;
#pragma unroll
  for (int j = J - 1; j >= 0; --j) A = (j == J - 1) ? vn[j] : fmaf(A, K.dd[j], vn[j]);
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    float dn = __shfl_down_sync(0xffffffffu, A, 1 << s);
    A = fmaf(K.bs[s], dn, A);
  }
  float d = __shfl_down_sync(0xffffffffu, A, 1);     // Ds at the first node of the next lane
  if (lane == 31) d = 0.f;

  // G(k) = vn(k) + Is(i+3,k) mi3(k) + Ds(k+1) md3(k);  Is(i,k) = Is(i+3,k) tII(k) + vn(k)
#pragma unroll
  for (int j = J - 1; j >= 0; --j) {
    float t = fmaf(S.I[S3][j], K.mi[j], vn[j]);
    X.g[j] = fmaf(d, K.md[j], t);
    d = fmaf(d, K.dd[j], vn[j]);
    S.I[S0][j] = fmaf(S.I[S3][j], K.ii[j], vn[j]);
  }
}

// the rest, from B(i): specials (:837-857) and Mt(i,.) = E(i) + G into slot S0; returns the row's specials, unscaled
template <int J, int PH>
__device__ __forceinline__ BckOut bck_row_back(BckState<J> &S, const BckMid<J> &X, float xB, const BckRowCtx &R)
{
  constexpr int S0 = PH, S3 = (PH + 3) & 3;
  float xC = S.xC[S3] * R.ploop;
  float xJ = fmaf(S.xJ[S3], R.ploop, xB * R.pmove);
  float xN = fmaf(S.xN[S3], R.ploop, xB * R.pmove);
  float xE = fmaf(xJ, R.tEL, xC * R.tEM);
#pragma unroll
  for (int j = 0; j < J; ++j) S.Mt[S0][j] = xE + X.g[j];
  return BckOut{ xB, xN, xJ, xC, xE };
}

template <int J, int VEC, int PH>
__device__ __forceinline__ BckOut bck_row_core(int lane, BckState<J> &S, const Bck3Consts<J> &K,
                                               const char *__restrict__ emis_lane, unsigned rowbytes, uint32_t cw, const BckRowCtx &R)
{
  BckMid<J> X;
  bck_row_front<J, VEC, PH>(lane, S, K, emis_lane, rowbytes, cw, X);
  return bck_row_back<J, PH>(S, X, warp_allsum(X.bsum), R);
}

// One Backward row.  PH = i & 3 (compile time).
template <int J, int VEC, int PH>
__device__ __forceinline__ void bck_row(int i, int lane, BckState<J> &S, const Bck3Consts<J> &K,
                                        const char *__restrict__ emis_lane, unsigned rowbytes, uint32_t cw,
                                        float fscale, BckRowCtx &R, float *__restrict__ xrow)
{
  constexpr int S0 = PH, S1 = (PH + 1) & 3, S2 = (PH + 2) & 3, S3 = (PH + 3) & 3;   // slots of rows i(=i+4), i+1, i+2, i+3
  const int L = R.L;

  if (i >= L - 1) {                 // warp-uniform: pad rows above L do nothing, rows L and L-1 initialise (:628-690)
    if (i <= L) {
      float xC = (i == L) ? R.pmove : R.ploop * R.pmove;
      float xE = xC * R.tEM;
      float sc = fscale;
      if (sc > 1.0f) {
        float sf = 1.0f / sc;
        xC *= sf; xE *= sf;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int j = 0; j < J; ++j) { S.Mt[r][j] *= sf; S.I[r][j] *= sf; }
        }
        R.totscale += logf(sc);
      }
#pragma unroll
      for (int j = 0; j < J; ++j) { S.Mt[S0][j] = xE; S.I[S0][j] = 0.f; }
      S.xN[S0] = 0.f; S.xJ[S0] = 0.f; S.xC[S0] = xC;
      if (i == L) S.xC[S1] = R.pmove;     // so that row L-2 reads C = tCL tCM (:815), whatever rows L, L-1 were scaled by
      if (lane == 0) {
        float2 *x2 = reinterpret_cast<float2 *>(xrow + (size_t)i * 6);
        x2[0] = make_float2(xE, 0.f);
        x2[1] = make_float2(0.f, 0.f);
        x2[2] = make_float2(xC, sc);
      }
    }
    return;
  }

  const BckOut O = bck_row_core<J, VEC, PH>(lane, S, K, emis_lane, rowbytes, cw, R);
  float xB = O.xB, xN = O.xN, xJ = O.xJ, xC = O.xC, xE = O.xE;

  if (i == 0) {                    // termination (:951-987): only B and N are defined on row 0, no rescaling
    S.xN[S0] = xN;
    if (lane == 0) {
      float2 *x2 = reinterpret_cast<float2 *>(xrow);
      x2[0] = make_float2(0.f, xN);
      x2[1] = make_float2(0.f, xB);
      x2[2] = make_float2(0.f, 1.0f);
    }
    return;
  }

  float scale = fscale;
  if (i < L - 2) {                 // (:910-916)
    if (xB > 1.0e16f) R.own_scales = true;
    if (R.own_scales) scale = (xB > 1.0e4f) ? xB : 1.0f;
  }
  if (scale > 1.0f) {              // warp-uniform
    float sf = 1.0f / scale;
    xN *= sf; xJ *= sf; xC *= sf; xB *= sf; xE *= sf;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < J; ++j) { S.Mt[r][j] *= sf; S.I[r][j] *= sf; }
      S.xN[r] *= sf; S.xJ[r] *= sf; S.xC[r] *= sf;
    }
    R.totscale += logf(scale);
  }
  S.xN[S0] = xN; S.xJ[S0] = xJ; S.xC[S0] = xC;
  if (lane == 0) {
    float2 *x2 = reinterpret_cast<float2 *>(xrow + (size_t)i * 6);
    x2[0] = make_float2(xE, xN);
    x2[1] = make_float2(xJ, xB);
    x2[2] = make_float2(xC, scale);
  }
}

// Rows i (phase PH) and i-1 (phase PH-1) in one basic block: row i-1 reads rows i+1 .. i+3 only, so the two rows are independent
// and the compiler interleaves their shuffle chains (the Forward parser's row-pair schedule, fs_parser_v3.cuh).  Everything is
// linear in the state, so row i's rescaling is applied afterwards to what row i-1 has just computed, then row i-1's own.
// Both rows must be regular rows: i <= L-2 and i-1 >= 1.
template <int J, int VEC, int PH>
__device__ __forceinline__ void bck_row_pair(int i, int lane, BckState<J> &S, const Bck3Consts<J> &K,
                                             const char *__restrict__ emis_lane, unsigned rowbytes, uint32_t cwA, uint32_t cwB,
                                             float fsA, float fsB, BckRowCtx &R, float *__restrict__ xrow)
{
  static_assert(PH == 3 || PH == 1, "pairs start on odd phases");
  BckOut A, B;
  if constexpr (BATHGPU_BCK_JOINT_B(J)) {            // both rows' B sums in one butterfly (pair_allsum, fs_parser.cuh)
    BckMid<J> XA, XB;
    bck_row_front<J, VEC, PH>(lane, S, K, emis_lane, rowbytes, cwA, XA);
    bck_row_front<J, VEC, PH - 1>(lane, S, K, emis_lane, rowbytes, cwB, XB);
    float xBA, xBB;
    pair_allsum(lane, XA.bsum, XB.bsum, xBA, xBB);
    A = bck_row_back<J, PH>(S, XA, xBA, R);
    B = bck_row_back<J, PH - 1>(S, XB, xBB, R);
  } else {
    A = bck_row_core<J, VEC, PH>(lane, S, K, emis_lane, rowbytes, cwA, R);
    B = bck_row_core<J, VEC, PH - 1>(lane, S, K, emis_lane, rowbytes, cwB, R);
  }
  const int L = R.L;
  auto rescale_all = [&](float sf) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < J; ++j) { S.Mt[r][j] *= sf; S.I[r][j] *= sf; }
      S.xN[r] *= sf; S.xJ[r] *= sf; S.xC[r] *= sf;
    }
  };
  float scaleA = fsA;
  if (i < L - 2) {                 // (:910-916)
    if (A.xB > 1.0e16f) R.own_scales = true;
    if (R.own_scales) scaleA = (A.xB > 1.0e4f) ? A.xB : 1.0f;
  }
  if (scaleA > 1.0f) {             // warp-uniform
    const float sf = 1.0f / scaleA;
    A.xN *= sf; A.xJ *= sf; A.xC *= sf; A.xB *= sf; A.xE *= sf;
    B.xN *= sf; B.xJ *= sf; B.xC *= sf; B.xB *= sf; B.xE *= sf;
    rescale_all(sf);
    R.totscale += logf(scaleA);
  }
  S.xN[PH] = A.xN; S.xJ[PH] = A.xJ; S.xC[PH] = A.xC;
  float scaleB = fsB;
  if (i - 1 < L - 2) {
    if (B.xB > 1.0e16f) R.own_scales = true;
    if (R.own_scales) scaleB = (B.xB > 1.0e4f) ? B.xB : 1.0f;
  }
  if (scaleB > 1.0f) {
    const float sf = 1.0f / scaleB;
    B.xN *= sf; B.xJ *= sf; B.xC *= sf; B.xB *= sf; B.xE *= sf;
    rescale_all(sf);
    R.totscale += logf(scaleB);
  }
  S.xN[PH - 1] = B.xN; S.xJ[PH - 1] = B.xJ; S.xC[PH - 1] = B.xC;
  if (lane == 0) {
    float2 *x2 = reinterpret_cast<float2 *>(xrow + (size_t)i * 6);
    x2[0] = make_float2(A.xE, A.xN);
    x2[1] = make_float2(A.xJ, A.xB);
    x2[2] = make_float2(A.xC, scaleA);
    x2 = reinterpret_cast<float2 *>(xrow + (size_t)(i - 1) * 6);
    x2[0] = make_float2(B.xE, B.xN);
    x2[1] = make_float2(B.xJ, B.xB);
    x2[2] = make_float2(B.xC, scaleB);
  }
}

template <int J> struct BckTune {
  static constexpr int kThreads   = 32;
#ifdef BATHGPU_BCK_WARPS
  static constexpr int kMinBlocks = BATHGPU_BCK_WARPS(J);
#else
  static constexpr int kMinBlocks = (J <= 3) ? 20 : (J == 4) ? 16 : (J == 5) ? 14 : (J == 6) ? 12 : (J == 7) ? 11 : (J == 8) ? 10 : 8;   // J = 6: 160 registers, no spill (+5 % over 128 with a 96-byte stack)
#endif
};

template <int J>
__global__ void __launch_bounds__(BckTune<J>::kThreads, BckTune<J>::kMinBlocks) fs3_backward_parser_kernel(FsBackwardArgs a)
{
  constexpr int VEC = VecOf<J>::V;
  const int lane = threadIdx.x & 31;

  Bck3Consts<J> K;
  load_bck3_consts<J>(a.cellb, lane, K);
  const char    *emis_lane = reinterpret_cast<const char *>(a.emis + lane * VEC);
  const unsigned rowbytes  = (unsigned)a.mpad * 4u;

  for (;;) {
    int w = 0;
    if (lane == 0) w = atomicAdd(a.counter, 1);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= a.nwin) break;

    const WindowDesc wd = a.wins[w];
    BckRowCtx R;
    R.L = wd.L; R.ploop = wd.ploop; R.pmove = wd.pmove; R.tEL = a.tEL; R.tEM = a.tEM;
    R.totscale = 0.f; R.own_scales = false;
    const float *fx   = a.fxmx + (size_t)a.xoff[w] * 6;
    float       *xrow = a.bxmx + (size_t)a.xoff[w] * 6;
    const int L = wd.L;

    BckState<J> S;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < J; ++j) { S.Mt[r][j] = 0.f; S.I[r][j] = 0.f; }
      S.xN[r] = 0.f; S.xJ[r] = 0.f; S.xC[r] = 0.f;
    }

    // rows 4*nq-1 (>= L) down to 0
    const int nq = (L + 4) >> 2;
    int i = 4 * nq - 1;
    for (int q0 = 0; q0 < nq; q0 += 8) {
      // lane l prepares row i - l: codon word from the nibbles n[i-l+1 .. i-l+4], and the Forward SCALE of that row
      const int myrow = i - lane;
      uint32_t cwl = 0;
      float    fsl = 1.0f;
      if (myrow >= 0) {
        long long nib = (wd.start - 1) + (long long)myrow + 8;       // 0-based block index of n[myrow+1], guard word included
        uint32_t lo = __ldg(a.dna4 + (nib >> 3)), hi = __ldg(a.dna4 + (nib >> 3) + 1);
        cwl = codon_word_bck(lo, hi, (int)(nib & 7) * 4, myrow, L);
        if (myrow <= L) fsl = __ldg(fx + (size_t)myrow * 6 + 5);
      }
      const int qn = min(8, nq - q0);
      for (int qq = 0; qq < qn; ++qq) {
#define BATHGPU_BPAIR(PH_)                                                                            \
        {                                                                                             \
          const int srcA = qq * 4 + (3 - PH_), srcB = srcA + 1;                                       \
          const uint32_t cwA = __shfl_sync(0xffffffffu, cwl, srcA), cwB = __shfl_sync(0xffffffffu, cwl, srcB); \
          const float    fsA = __shfl_sync(0xffffffffu, fsl, srcA), fsB = __shfl_sync(0xffffffffu, fsl, srcB); \
          if (i <= L - 2 && i >= 2) bck_row_pair<J, VEC, PH_>(i, lane, S, K, emis_lane, rowbytes, cwA, cwB, fsA, fsB, R, xrow); \
          else {                                                                                      \
            bck_row<J, VEC, PH_>(i, lane, S, K, emis_lane, rowbytes, cwA, fsA, R, xrow);              \
            bck_row<J, VEC, PH_ - 1>(i - 1, lane, S, K, emis_lane, rowbytes, cwB, fsB, R, xrow);      \
          }                                                                                           \
          i -= 2;                                                                                     \
        }
        BATHGPU_BPAIR(3) BATHGPU_BPAIR(1)
#undef BATHGPU_BPAIR
      }
    }

    // score (:989-1003): N(0) + N(1) + N(2), rows 1 and 2 as left by every later rescale
    {
      float tot = S.xN[0] + S.xN[1] + S.xN[2];
      int   st  = 0;
      float sc;
      if (isnan(tot) || isinf(tot)) { st = 16; sc = tot; }
      else if (tot == 0.0f)         { st = 16; sc = -INFINITY; }
      else sc = R.totscale + logf(tot);
      if (lane == 0) { a.bcksc[w] = sc; if (st) a.status[w] = st; }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// p7_FLogsum (src/logsum.c:104-111): table-driven log(e^a + e^b); the table entry is recomputed
// in double exactly as p7_FLogsumInit fills it (:80-91).
__device__ __forceinline__ float flogsum_dev(float a, float b)
{
  const float mx = (a > b) ? a : b;
  const float mn = (a > b) ? b : a;
  if (mn == -INFINITY || (mx - mn) >= 15.7f) return mx;
  const int idx = (int)((mx - mn) * 1000.f);
  return mx + (float)log(1. + exp((double)-idx / 1000.));
}

struct DomainDecodeArgs {
  const float     *fxmx, *bxmx;      // X rows of both parsers, window w at xoff[w]*6
  const long long *xoff;
  const WindowDesc *wins;
  int              nwin;
  float            tNL, tJL, tCL;    // {N,J,C}->LOOP odds of the profile the caller passes (om_fs5, src/p7_domaindef.c:320)
  float           *lsf, *lsb;        // scratch: cumulative log scales, window w at xoff[w] (+1 slot at the end for lsb)
  float           *mocc, *btot, *etot;     // out, window w at ooff[w], L+1 floats each
  const long long *ooff;
  const int       *status;           // windows whose parsers failed are skipped (outputs zeroed)
};

// One warp per window.  The cumulative sums run serially in one lane in the reference's order (the
// terms are exact zeros except at the few rescaled rows); everything with an expf in it runs lane-parallel.
static __global__ void __launch_bounds__(128) fs_domain_decoding_kernel(DomainDecodeArgs a)
{
  const int lane = threadIdx.x & 31;
  const int w    = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= a.nwin) return;
  const int L = a.wins[w].L;
  const float *xf = a.fxmx + (size_t)a.xoff[w] * 6;
  const float *xb = a.bxmx + (size_t)a.xoff[w] * 6;
  float *lsf = a.lsf + (size_t)a.xoff[w] + 2 * (size_t)w;      // L+2 slots per window
  float *lsb = a.lsb + (size_t)a.xoff[w] + 2 * (size_t)w;
  float *mocc = a.mocc + a.ooff[w], *btot = a.btot + a.ooff[w], *etot = a.etot + a.ooff[w];

  if (a.status[w] != 0) {
    for (int i = lane; i <= L; i += 32) { mocc[i] = 0.f; btot[i] = 0.f; etot[i] = 0.f; }
    return;
  }

  for (int i = lane; i <= L; i += 32) { lsf[i] = logf(xf[i * 6 + 5]); lsb[i] = logf(xb[i * 6 + 5]); }
  __syncwarp();
  if (lane == 0) {                 // log_sfwd[i] = sum_{0..i}, :262-264
    float acc = lsf[0];
    for (int i = 1; i <= L; ++i) { acc = acc + lsf[i]; lsf[i] = acc; }
  } else if (lane == 1) {          // log_sbck[i] = sum_{i..L}, :269-271
    float acc = 0.0f;
    lsb[L + 1] = 0.0f;
    for (int i = L; i >= 0; --i) { acc = acc + lsb[i]; lsb[i] = acc; }
  }
  __syncwarp();

  const float liz = -flogsum_dev(logf(xb[0 * 6 + 1]) + lsb[0],
                                 flogsum_dev(logf(xb[1 * 6 + 1]) + lsb[1], logf(xb[2 * 6 + 1]) + lsb[2]));   // :277-282

  // per-row terms: btot/etot increments parked in the output arrays, mocc final           (:296-352)
  for (int i = lane; i <= L; i += 32) {
    if (i < 3) { mocc[i] = 0.f; btot[i] = 0.f; etot[i] = 0.f; continue; }
    btot[i] = xf[(i - 3) * 6 + 3] * xb[(i - 3) * 6 + 3] * expf(lsf[i - 3] + lsb[i - 3] + liz);
    etot[i] = xf[i * 6 + 0] * xb[i * 6 + 0] * expf(lsf[i] + lsb[i] + liz);
    float njcp = 0.f;
#pragma unroll
    for (int s = 0; s < 3; ++s) {          // s: 0 = N (cell 1), 1 = J (cell 2), 2 = C (cell 4)
      const int   cell = (s == 0) ? 1 : (s == 1) ? 2 : 4;
      const float tl   = (s == 0) ? a.tNL : (s == 1) ? a.tJL : a.tCL;
      njcp += xf[(i - 3) * 6 + cell] * xb[i * 6 + cell] * tl * expf(lsf[i - 3] + lsb[i] + liz);
      if (i < L)     njcp += xf[(i - 2) * 6 + cell] * xb[(i + 1) * 6 + cell] * tl * expf(lsf[i - 2] + lsb[i + 1] + liz);
      if (i < L - 1) njcp += xf[(i - 1) * 6 + cell] * xb[(i + 2) * 6 + cell] * tl * expf(lsf[i - 1] + lsb[i + 2] + liz);
    }
    mocc[i] = 1.f - njcp;
  }
  __syncwarp();
  if (lane < 3) {                  // three interleaved running sums, stride 3 (:299,305)
    float bacc = 0.f, eacc = 0.f;
    for (int i = 3 + lane; i <= L; i += 3) {
      bacc = bacc + btot[i]; btot[i] = bacc;
      eacc = eacc + etot[i]; etot[i] = eacc;
    }
  }
}

}  // namespace bathgpu
