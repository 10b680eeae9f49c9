// orf_domain.cuh -- the standard-translation branch's DP over ORFs (amino-acid sequences) for sm_100a:
//   p7_Forward / p7_ForwardParser, p7_Backward / p7_BackwardParser (reference src/impl_sse/fwdback.c:256-466, :468-738),
//   p7_Decoding (decoding.c:76-139), p7_Null2_ByExpectation (null2.c:44-125), p7_OptimalAccuracy (optacc.c:58-174)
//   and p7_OATrace (:225-425), as called from p7_pli_Frameshift's standard branch (src/p7_pipeline.c:1480-1511) and
//   rescore_isolated_domain_bath (src/p7_domaindef.c:1229-1370).
//
// Same decomposition as fs_domain.cuh, of which this is the one-residue-look-back case: one warp per ORF / envelope,
// lane l owns J contiguous nodes, rows sequential, the emission table is the amino-acid part (rows amino0 + x) of the
// loaded frameshift table -- already multiplied by tBM(k-1) Z(k), so E(i) = sum of the match cells and the Backward
// sweep carries M/Z -- and the lane constants are the ones the 5-codon kernels use.
//   orf_forward_kernel<J,FULL>   X rows (and, FULL, the cells {I, M Z} per node);
//   orf_backward_kernel<J,FULL>  X rows; FULL: posterior decoding of each row as it is produced, written over the
//                                Forward cells, null2 sums in registers.  The reference normalises with Backward's
//                                N(0), known only after the sweep; this kernel uses the Forward score, the same
//                                path sum (they agree to 1e-5 relative), which is what lets decoding run inside the sweep;
//   orf_optacc_kernel<J>, orf_oatrace_kernel: max-plus fill and the traceback state machine.
#pragma once
#include "fs_domain.cuh"

namespace bathgpu {

constexpr int kPPCellsP = 2;          // per node and row: I, M
enum PPCellP { PPP_I = 0, PPP_M = 1 };

struct OrfDomainArgs {
  const float    *emis;        // folded table [nrows][mpad]
  int             amino0;      // first amino-acid row of that table (338 or 1367)
  const float    *amino;       // unfolded amino-acid odds [20][mpad] (null2)
  const float    *cellf;       // Fwd5Consts image
  const float    *cellb;       // BckConsts image
  const float    *oapass;
  const uint32_t *oaflags;
  const uint8_t  *residues;
  const EnvelopeDesc *envs;    // start = 0-based offset of the first residue in the residue buffer
  int             nenv;
  int             M, mpad;
  float           tEM, tEL;
  const long long *xoff;
  float          *pp;          // [rows][2][mpad]
  float          *dcell;       // [rows][mpad]  Forward D cells (orf_forward_kernel<J, true, true> only: the stochastic trace reads them)
  float          *oa;          // [rows][3][mpad]
  float          *fx, *bx;     // X rows {E,N,J,B,C,SCALE}
  float          *ppx, *oax;
  float          *lsf;
  float          *fwdsc, *bcksc, *oasc;
  float          *null2;
  int            *status;
  int            *counter;
  int             lanes_f32;   // floats per SIMD vector of the CPU build whose E-state tie-break is reproduced (4: SSE)
};

template <int J, bool FULL, bool KEEPD = false>
__global__ void __launch_bounds__(32) orf_forward_kernel(OrfDomainArgs a)
{
  constexpr int VEC = VecOf<J>::V;
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  Fwd5Consts<J> K;
  load_fwd5_consts<J>(a.cellf, lane, K);
  const float *emis_lane = a.emis + (size_t)a.amino0 * a.mpad + lane * VEC;

  for (;;) {
    int e = 0;
    if (lane == 0) e = atomicAdd(a.counter, 1);
    e = __shfl_sync(full, e, 0);
    if (e >= a.nenv) break;
    const EnvelopeDesc ed = a.envs[e];
    const int L = ed.L;
    const float pmove = ed.pmove, ploop = ed.ploop;
    const long long xo = a.xoff[e];
    float *fxrow = a.fx + (size_t)xo * 6;
    float *lsfrow = a.lsf + xo;
    float *pprow_lane = FULL ? a.pp + (size_t)xo * kPPCellsP * a.mpad + lane * VEC : nullptr;
    float *drow_lane = KEEPD ? a.dcell + (size_t)xo * a.mpad + lane * VEC : nullptr;

    float W[J], I[J];
#pragma unroll
    for (int j = 0; j < J; ++j) { W[j] = pmove; I[j] = 0.f; }          // row 0: B = tNB (:281-285)
    float xN = 1.0f, xJ = 0.f, xC = 0.f, totscale = 0.f, lsf = 0.f;
    if (lane == 0) {
      fxrow[0] = 0.f; fxrow[1] = 1.0f; fxrow[2] = 0.f; fxrow[3] = pmove; fxrow[4] = 0.f; fxrow[5] = 1.0f;
      lsfrow[0] = 0.f;
    }
    if constexpr (FULL) {
      float z[J];
#pragma unroll
      for (int j = 0; j < J; ++j) z[j] = 0.f;
      store_row<J, VEC>(pprow_lane + PPP_I * a.mpad, z);
      store_row<J, VEC>(pprow_lane + PPP_M * a.mpad, z);
      if constexpr (KEEPD) store_row<J, VEC>(drow_lane, z);
    }

    int chunk = -64;
    unsigned myres = 0;
    for (int i = 1; i <= L; ++i) {
      if (i >= chunk + 32) { chunk = i; myres = (i + lane <= L) ? a.residues[ed.start + i + lane - 1] : 0u; }
      const unsigned x = __shfl_sync(full, myres, i - chunk);
      float ev[J], m[J], icur[J];
      load_emission_row<J, VEC>(emis_lane + (size_t)x * a.mpad, ev);
      float es = 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) { m[j] = W[j] * ev[j]; es += m[j]; icur[j] = I[j]; }
      float xE = warp_allsum(es);

      float av[J];
      float A = 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) { av[j] = m[j] * K.md[j]; A = (j == 0) ? av[0] : fmaf(A, K.dd[j], av[j]); }
#pragma unroll
      for (int s = 0; s < 5; ++s) {
        float up = __shfl_up_sync(full, A, 1 << s);
        A = fmaf(K.bs[s], up, A);
      }
      float d = __shfl_up_sync(full, A, 1);
      if (lane == 0) d = 0.f;

      xN = xN * ploop;                                     // (:401-404)
      xC = fmaf(xC, ploop, xE * a.tEM);
      xJ = fmaf(xJ, ploop, xE * a.tEL);
      float xB = fmaf(xJ, pmove, xN * pmove);

      float ov[J];
      float dv[KEEPD ? J : 1];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        float t = fmaf(icur[j], K.im[j], m[j] * K.mm[j]);
        ov[j] = fmaf(d, K.dm[j], t);
        if constexpr (KEEPD) dv[j] = d;
        if (j + 1 < J) d = fmaf(d, K.dd[j], av[j]);
        I[j] = fmaf(icur[j], K.ii[j], m[j] * K.mi[j]);     // I(i+1,k)
      }
      float oprev = __shfl_up_sync(full, ov[J - 1], 1);
      if (lane == 0) oprev = 0.f;
      W[0] = xB + oprev;
#pragma unroll
      for (int j = 1; j < J; ++j) W[j] = xB + ov[j - 1];

      float scale = 1.0f;
      if (xE > 1.0e4f) {                                   // sparse rescaling (:407-423): row i and everything derived from it
        const float sf = 1.0f / xE;
        scale = xE;
        xN *= sf; xC *= sf; xJ *= sf; xB *= sf;
#pragma unroll
        for (int j = 0; j < J; ++j) { W[j] *= sf; I[j] *= sf; m[j] *= sf; icur[j] *= sf; if constexpr (KEEPD) dv[j] *= sf; }
        totscale += logf(xE);
        xE = 1.0f;
      }
      lsf += logf(scale);
      if constexpr (FULL) {
        float *row = pprow_lane + (size_t)i * kPPCellsP * a.mpad;
        store_row<J, VEC>(row + PPP_I * a.mpad, icur);
        store_row<J, VEC>(row + PPP_M * a.mpad, m);
        if constexpr (KEEPD) store_row<J, VEC>(drow_lane + (size_t)i * a.mpad, dv);
      }
      if (lane == 0) {
        float2 *x2 = reinterpret_cast<float2 *>(fxrow + (size_t)i * 6);
        x2[0] = make_float2(xE, xN);
        x2[1] = make_float2(xJ, xB);
        x2[2] = make_float2(xC, scale);
        lsfrow[i] = lsf;
      }
    }
    int   st = 0;
    float sc;
    if (isnan(xC))                 { st = 16; sc = xC; }                 // (:447-449)
    else if (L > 0 && xC == 0.0f)  { st = 16; sc = -INFINITY; }
    else if (isinf(xC))            { st = 16; sc = xC; }
    else sc = totscale + logf(xC * pmove);
    if (lane == 0) { a.fwdsc[e] = sc; a.status[e] = st; if (FULL) a.oasc[e] = xC * pmove; }   // oasc[] carries the unscaled path sum to the Backward sweep
  }
}

// Backward, rows L down to 0.  Carries Mt(i,k) = M(i,k) / Z(k) (see fs_backward.cuh): Mt(i,k) = E(i) + G(i,k).
template <int J, bool FULL>
__global__ void __launch_bounds__(32) orf_backward_kernel(OrfDomainArgs a)
{
  constexpr int VEC = VecOf<J>::V;
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  BckConsts<J> K;
  load_bck_consts<J>(a.cellb, lane, K);
  const float *emis_lane = a.emis + (size_t)a.amino0 * a.mpad + lane * VEC;

  for (;;) {
    int e = 0;
    if (lane == 0) e = atomicAdd(a.counter, 1);
    e = __shfl_sync(full, e, 0);
    if (e >= a.nenv) break;
    if (a.status[e] != 0) { if (lane == 0) a.bcksc[e] = -INFINITY; continue; }
    const EnvelopeDesc ed = a.envs[e];
    const int L = ed.L;
    const float pmove = ed.pmove, ploop = ed.ploop;
    const long long xo = a.xoff[e];
    const float *fxrow = a.fx + (size_t)xo * 6;
    const float *lsfrow = a.lsf + xo;
    float *bxrow = a.bx + (size_t)xo * 6;
    float *pprow_lane = FULL ? a.pp + (size_t)xo * kPPCellsP * a.mpad + lane * VEC : nullptr;
    float *ppxrow = FULL ? a.ppx + (size_t)xo * 6 : nullptr;
    const float liz = -a.fwdsc[e];
    // 1 / (Forward's path sum in Forward's own scaling): what 1 / bck N(0) is while Backward runs on Forward's scale factors
    const float invz = FULL ? 1.0f / a.oasc[e] : 0.f;

    float Mt[J], I[J], accM[J], accI[J];
    float accN = 0.f, accJ = 0.f, accC = 0.f;
    float xN = 0.f, xJ = 0.f, xB = 0.f, xC = pmove, xE = pmove * a.tEM;      // row L (:487-492)
    float totscale = 0.f, lsb = 0.f;
    bool  own = false;
    int   st = 0;
#pragma unroll
    for (int j = 0; j < J; ++j) { Mt[j] = xE; I[j] = 0.f; accM[j] = 0.f; accI[j] = 0.f; }

    int chunk = 1 << 30;
    unsigned myres = 0;
    for (int i = L; i >= 0; --i) {
      if constexpr (FULL) { if (i >= 1) prefetch_matrix_row(pprow_lane - lane * VEC + (size_t)(i - 1) * kPPCellsP * a.mpad, kPPCellsP, a.mpad, lane); }
      float scale = 1.0f;
      if (i < L) {
        // residue x_{i+1}
        if (i + 1 < chunk) { chunk = i + 1 - 31; myres = (chunk + lane >= 1) ? a.residues[ed.start + chunk + lane - 1] : 0u; }
        const unsigned x = __shfl_sync(full, myres, i + 1 - chunk);
        float ev[J], v[J];
        load_emission_row<J, VEC>(emis_lane + (size_t)x * a.mpad, ev);
        float bsum = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) { v[j] = Mt[j] * ev[j]; bsum += v[j]; }
        xB = warp_allsum(bsum);
        if (i == 0) {                                      // termination (:697-720)
          xN = fmaf(xB, pmove, xN * ploop);
          if (lane == 0) { bxrow[0] = 0.f; bxrow[1] = xN; bxrow[2] = 0.f; bxrow[3] = xB; bxrow[4] = 0.f; bxrow[5] = 1.0f; }
          break;
        }
        float vn[J];
        {
          float up = __shfl_down_sync(full, v[0], 1);
          if (lane == 31) up = 0.f;
#pragma unroll
          for (int j = 0; j + 1 < J; ++j) vn[j] = v[j + 1];
          vn[J - 1] = up;
        }
        float av[J];
        float A = 0.f;
#pragma unroll
        for (int j = J - 1; j >= 0; --j) { av[j] = vn[j] * K.vdm[j]; A = (j == J - 1) ? av[j] : fmaf(A, K.dd[j], av[j]); }
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          float dn = __shfl_down_sync(full, A, 1 << s);
          A = fmaf(K.bs[s], dn, A);
        }
        float d = __shfl_down_sync(full, A, 1);
        if (lane == 31) d = 0.f;

        xC = xC * ploop;                                   // (:599-603)
        xJ = fmaf(xB, pmove, xJ * ploop);
        xN = fmaf(xB, pmove, xN * ploop);
        xE = fmaf(xC, a.tEM, xJ * a.tEL);
#pragma unroll
        for (int j = J - 1; j >= 0; --j) {
          float t = I[j] * K.mi[j];
          t = fmaf(vn[j], K.vmm[j], t);
          float g = fmaf(d, K.md[j], t);
          d = fmaf(d, K.dd[j], av[j]);
          I[j] = fmaf(I[j], K.ii[j], vn[j] * K.vim[j]);
          Mt[j] = xE + g;
        }
        if (xB > 1.0e16f) own = true;                      // (:657-660)
        scale = own ? ((xB > 1.0e4f) ? xB : 1.0f) : fxrow[(size_t)i * 6 + 5];
      } else scale = fxrow[(size_t)L * 6 + 5];
      if (scale > 1.0f) {
        const float sf = 1.0f / scale;
        xE *= sf; xN *= sf; xJ *= sf; xB *= sf; xC *= sf;
#pragma unroll
        for (int j = 0; j < J; ++j) { Mt[j] *= sf; I[j] *= sf; }
        totscale += logf(scale);
      }
      lsb += logf(scale);
      if (lane == 0) {
        float2 *x2 = reinterpret_cast<float2 *>(bxrow + (size_t)i * 6);
        x2[0] = make_float2(xE, xN);
        x2[1] = make_float2(xJ, xB);
        x2[2] = make_float2(xC, scale);
      }
      if constexpr (FULL) {                                // decoding of row i (decoding.c:104-131)
        float *row = pprow_lane + (size_t)i * kPPCellsP * a.mpad;
        float fI[J], fM[J];
        load_row<J, VEC>(row + PPP_I * a.mpad, fI);
        load_row<J, VEC>(row + PPP_M * a.mpad, fM);
        // posterior = fwd * bck * scale(i) / Z (decoding.c:107-121); with Backward's own scale factors in play the
        // cumulative log scales of both sweeps are needed instead
        const float fac = own ? expf(lsfrow[i] + lsb + liz) : fxrow[(size_t)i * 6 + 5] * invz;
        if (isinf(fac)) st = 16;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          fI[j] = (fI[j] * I[j]) * fac; fM[j] = (fM[j] * Mt[j]) * fac;
          accM[j] += fM[j]; accI[j] += fI[j];
        }
        store_row<J, VEC>(row + PPP_I * a.mpad, fI);
        store_row<J, VEC>(row + PPP_M * a.mpad, fM);
        const float facx = (own ? expf(lsfrow[i - 1] + lsb + liz) : invz) * ploop;
        const float *f1 = fxrow + (size_t)(i - 1) * 6;
        const float pN = f1[1] * xN * facx, pJ = f1[2] * xJ * facx, pC = f1[4] * xC * facx;
        accN += pN; accJ += pJ; accC += pC;
        if (lane == 0) {
          float2 *x2 = reinterpret_cast<float2 *>(ppxrow + (size_t)i * 6);
          x2[0] = make_float2(0.f, pN);
          x2[1] = make_float2(pJ, 0.f);
          x2[2] = make_float2(pC, scale);
        }
      }
    }
    {
      float sc;
      if (isnan(xN) || isinf(xN))      { st = 16; sc = xN; }
      else if (L > 0 && xN == 0.0f)    { st = 16; sc = -INFINITY; }
      else sc = totscale + logf(xN);
      if (lane == 0) { a.bcksc[e] = sc; if (st) a.status[e] = st; }
    }
    if constexpr (FULL) {
      if (lane == 0) { for (int s = 0; s < 6; ++s) ppxrow[s] = 0.f; }
      // null2 by expectation (null2.c:57-118)
      const float norm = 1.0f / (float)L;
      const float xfactor = accN * norm + accC * norm + accJ * norm;
      float isum = 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) { accM[j] *= norm; accI[j] *= norm; isum += accI[j]; }
      isum = warp_allsum(isum);
      float n2 = 0.f;
      for (int x = 0; x < 20; ++x) {
        float r[J];
        load_emission_row<J, VEC>(a.amino + (size_t)x * a.mpad + lane * VEC, r);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) s = fmaf(accM[j], r[j], s);
        s = warp_allsum(s) + isum + xfactor;
        if (lane == x) n2 = s;
      }
      float out = 1.0f;
      float vals[20];
#pragma unroll
      for (int x = 0; x < 20; ++x) vals[x] = __shfl_sync(full, n2, x);
      if (lane < 20) out = n2;
      else if (lane == 21) out = (vals[2] + vals[11]) / 2.0f;      // B = D,N   (esl_abc_FAvgScVec)
      else if (lane == 22) out = (vals[7] + vals[9]) / 2.0f;       // J = I,L
      else if (lane == 23) out = (vals[3] + vals[13]) / 2.0f;      // Z = E,Q
      else if (lane == 24) out = vals[8];                          // O = K
      else if (lane == 25) out = vals[1];                          // U = C
      else if (lane == 26) { float s = 0.f;
#pragma unroll
        for (int x = 0; x < 20; ++x) s += vals[x];
        out = s / 20.0f; }
      if (lane < 29) a.null2[(size_t)e * 29 + lane] = out;
    }
  }
}

// Optimal-accuracy fill (optacc.c:58-174): max-plus, transitions are masks, a forbidden path contributes 0.0.
template <int J>
__global__ void __launch_bounds__(32) orf_optacc_kernel(OrfDomainArgs a)
{
  constexpr int VEC = VecOf<J>::V;
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const int J0 = lane * J;
  uint32_t fl[J];
#pragma unroll
  for (int j = 0; j < J; ++j) fl[j] = a.oaflags[J0 + j];
  float dpass[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) dpass[s] = a.oapass[s * 32 + lane];
  const int M = a.M;

  for (;;) {
    int e = 0;
    if (lane == 0) e = atomicAdd(a.counter, 1);
    e = __shfl_sync(full, e, 0);
    if (e >= a.nenv) break;
    if (a.status[e] != 0) { if (lane == 0) a.oasc[e] = -INFINITY; continue; }
    const EnvelopeDesc ed = a.envs[e];
    const int L = ed.L;
    const long long xo = a.xoff[e];
    const float *pprow_lane = a.pp + (size_t)xo * kPPCellsP * a.mpad + lane * VEC;
    float *oarow_lane = a.oa + (size_t)xo * kOACells * a.mpad + lane * VEC;
    const float *ppxrow = a.ppx + (size_t)xo * 6;
    float *oaxrow = a.oax + (size_t)xo * 6;
    const bool loopN = ed.ploop != 0.f, loopJ = loopN, loopC = loopN;
    const bool moveN = ed.pmove != 0.f, moveJ = moveN;
    const bool loopE = a.tEL != 0.f, moveE = a.tEM != 0.f;

    float P[J], Mr[J], Ir[J];
    float xN = 0.f, xJ = -INFINITY, xC = -INFINITY;
#pragma unroll
    for (int j = 0; j < J; ++j) {                         // row 0 (:87-94): cells -inf, B = 0
      float s = oa_mask(fl[j], OF_BM, 0.0f);
      s = oa_max(s, oa_mask(fl[j], OF_MM, -INFINITY));
      s = oa_max(s, oa_mask(fl[j], OF_IM, -INFINITY));
      s = oa_max(s, oa_mask(fl[j], OF_DM, -INFINITY));
      P[j] = s; Mr[j] = -INFINITY; Ir[j] = -INFINITY;
    }
    {
      float ninf[J];
#pragma unroll
      for (int j = 0; j < J; ++j) ninf[j] = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < kOACells; ++cc) store_row<J, VEC>(oarow_lane + (size_t)cc * a.mpad, ninf);
      if (lane == 0) { oaxrow[0] = -INFINITY; oaxrow[1] = 0.f; oaxrow[2] = -INFINITY; oaxrow[3] = 0.f; oaxrow[4] = -INFINITY; oaxrow[5] = 0.f; }
    }
    for (int i = 1; i <= L; ++i) {
      const float *row = pprow_lane + (size_t)i * kPPCellsP * a.mpad;
      if (i < L) prefetch_matrix_row(row - lane * VEC + (size_t)kPPCellsP * a.mpad, kPPCellsP, a.mpad, lane);
      float pc[J], mnew[J], inew[J], dnew[J];
      load_row<J, VEC>(row + PPP_M * a.mpad, pc);
#pragma unroll
      for (int j = 0; j < J; ++j) mnew[j] = P[j] + pc[j];
      load_row<J, VEC>(row + PPP_I * a.mpad, pc);
#pragma unroll
      for (int j = 0; j < J; ++j) {
        float s = oa_mask(fl[j], OF_MI, Mr[j]);
        s = oa_max(s, oa_mask(fl[j], OF_II, Ir[j]));
        inew[j] = s + pc[j];
        if (J0 + j + 1 > M) { inew[j] = -INFINITY; mnew[j] = -INFINITY; }
      }
      {   // D(i,k) = max(mask(MD(k-1)) M(i,k-1), mask(DD(k-1)) D(i,k-1)), D(i,1) = -inf (:127-151)
        float mprev = __shfl_up_sync(full, mnew[J - 1], 1);
        if (lane == 0) mprev = -INFINITY;
        float endv;
        {
          float dd_ = -INFINITY, mm_ = mprev;
#pragma unroll
          for (int j = 0; j < J; ++j) {
            float dv = oa_max(oa_mask(fl[j], OF_DD, dd_), oa_mask(fl[j], OF_MD, mm_));
            if (J0 + j + 1 == 1) dv = -INFINITY;
            dd_ = dv; mm_ = mnew[j];
          }
          endv = dd_;
        }
        float A = endv;
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          float up = __shfl_up_sync(full, A, 1 << s);
          if (lane >= (1 << s) && dpass[s] != 0.f) A = oa_max(A, up);
        }
        float carry = __shfl_up_sync(full, A, 1);
        if (lane == 0) carry = -INFINITY;
        float dd_ = carry, mm_ = mprev;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          float dv = oa_max(oa_mask(fl[j], OF_DD, dd_), oa_mask(fl[j], OF_MD, mm_));
          const int k = J0 + j + 1;
          if (k == 1 || k > M) dv = -INFINITY;
          dnew[j] = dv;
          dd_ = dv; mm_ = mnew[j];
        }
      }
      float ee = -INFINITY;
#pragma unroll
      for (int j = 0; j < J; ++j) ee = oa_max(ee, oa_max(mnew[j], dnew[j]));
#pragma unroll
      for (int dlt = 16; dlt >= 1; dlt >>= 1) ee = oa_max(ee, __shfl_xor_sync(full, ee, dlt));
      const float xE = ee;

      const float ppN = ppxrow[(size_t)i * 6 + 1], ppJ = ppxrow[(size_t)i * 6 + 2], ppC = ppxrow[(size_t)i * 6 + 4];
      float t1 = loopJ ? xJ + ppJ : 0.0f, t2 = loopE ? xE : 0.0f;      // (:157-170)
      xJ = (t1 > t2) ? t1 : t2;
      t1 = loopC ? xC + ppC : 0.0f; t2 = moveE ? xE : 0.0f;
      xC = (t1 > t2) ? t1 : t2;
      xN = loopN ? xN + ppN : 0.0f;
      t1 = moveN ? xN : 0.0f; t2 = moveJ ? xJ : 0.0f;
      const float xB = (t1 > t2) ? t1 : t2;
      {
        float mp = __shfl_up_sync(full, mnew[J - 1], 1);
        float ip = __shfl_up_sync(full, inew[J - 1], 1);
        float dp = __shfl_up_sync(full, dnew[J - 1], 1);
        if (lane == 0) { mp = -INFINITY; ip = -INFINITY; dp = -INFINITY; }
#pragma unroll
        for (int j = 0; j < J; ++j) {
          float s = oa_mask(fl[j], OF_BM, xB);
          s = oa_max(s, oa_mask(fl[j], OF_MM, mp));
          s = oa_max(s, oa_mask(fl[j], OF_IM, ip));
          s = oa_max(s, oa_mask(fl[j], OF_DM, dp));
          P[j] = s;
          mp = mnew[j]; ip = inew[j]; dp = dnew[j];
        }
      }
#pragma unroll
      for (int j = 0; j < J; ++j) { Mr[j] = mnew[j]; Ir[j] = inew[j]; }
      float *orow = oarow_lane + (size_t)i * kOACells * a.mpad;
      store_row<J, VEC>(orow + OA_M * a.mpad, mnew);
      store_row<J, VEC>(orow + OA_I * a.mpad, inew);
      store_row<J, VEC>(orow + OA_D * a.mpad, dnew);
      if (lane == 0) {
        float2 *x2 = reinterpret_cast<float2 *>(oaxrow + (size_t)i * 6);
        x2[0] = make_float2(xE, xN);
        x2[1] = make_float2(xJ, xB);
        x2[2] = make_float2(xC, 0.f);
      }
    }
    if (lane == 0) a.oasc[e] = xC;                       // ret_e = C(L) (:172)
  }
}

// Traceback (optacc.c:225-425).  One warp per envelope, the lanes split the E-state argmax.
static __global__ void __launch_bounds__(128) orf_oatrace_kernel(OrfDomainArgs a, TraceArgs t)
{
  const int lane = threadIdx.x & 31;
  const int e    = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= a.nenv) return;
  if (a.status[e] != 0) { if (lane == 0) t.tlen[e] = 0; return; }
  const EnvelopeDesc ed = a.envs[e];
  const int L = ed.L, M = a.M, mpad = a.mpad, J = t.J;
  const long long xo = a.xoff[e];
  const float *pp  = a.pp + (size_t)xo * kPPCellsP * mpad;
  const float *oa  = a.oa + (size_t)xo * kOACells * mpad;
  const float *ppx = a.ppx + (size_t)xo * 6;
  const float *oax = a.oax + (size_t)xo * 6;
  TraceStep *out = t.steps + t.toff[e];
  const int ld = M + 1;
  auto TF = [&](int tt, int k) -> float { return t.tfv[(size_t)tt * ld + k]; };
  auto OA = [&](int i, int k, int cell) -> float {
    if (k < 1) return 0.0f;                                // rightshiftz shifts in 0.0 (:295-299); node 0's transitions are 0 anyway
    return oa[((size_t)i * kOACells + cell) * mpad + perm_of(k, J)];
  };
  auto PP = [&](int i, int k, int cell) -> float { return pp[((size_t)i * kPPCellsP + cell) * mpad + perm_of(k, J)]; };
  const bool loopC = ed.ploop != 0.f, loopJ = loopC, moveN = ed.pmove != 0.f, moveJ = moveN;
  const bool loopE = a.tEL != 0.f, moveE = a.tEM != 0.f;
  int Q = (M - 1) / a.lanes_f32 + 1;
  if (Q < 2) Q = 2;

  int n = 0, last_st = 0;
  auto emit = [&](int st, int k, int i, float p) {         // p7_trace_AppendWithPP: which fields each state keeps
    TraceStep s; s.st = (uint8_t)st; s.i = 0; s.k = 0; s.c = 0; s.pp = 0.f;
    if (st == TS_N || st == TS_C || st == TS_J) { if (last_st == st) { s.i = i; s.pp = p; } }
    else if (st == TS_D) s.k = (int16_t)k;
    else if (st == TS_M || st == TS_I) { s.i = i; s.k = (int16_t)k; s.pp = p; }
    if (lane == 0) out[n] = s;
    last_st = st;
    ++n;
  };
  int i = L, k = 0;
  emit(TS_T, k, i, 0.f);
  emit(TS_C, k, i, 0.f);
  int s0 = TS_C, s1 = TS_C;
  const int max_steps = L + M + 8;
  bool bad = false;
  while (s0 != TS_S) {
    switch (s0) {
    case TS_M: {                                           // select_m (:283-313): order M > I > D > B
      float pm = (TF(1, k - 1) == 0.f) ? -INFINITY : OA(i - 1, k - 1, OA_M);
      float pi = (TF(2, k - 1) == 0.f) ? -INFINITY : OA(i - 1, k - 1, OA_I);
      float pd = (TF(3, k - 1) == 0.f) ? -INFINITY : OA(i - 1, k - 1, OA_D);
      float pb = (TF(0, k - 1) == 0.f) ? -INFINITY : oax[(size_t)(i - 1) * 6 + 3];
      s1 = TS_M; float best = pm;
      if (pi > best) { best = pi; s1 = TS_I; }
      if (pd > best) { best = pd; s1 = TS_D; }
      if (pb > best) { best = pb; s1 = TS_B; }
      k--; i--; break; }
    case TS_D: {                                           // select_d (:317-342)
      float pm = (TF(4, k - 1) == 0.f) ? -INFINITY : OA(i, k - 1, OA_M);
      float pd = (TF(7, k - 1) == 0.f) ? -INFINITY : OA(i, k - 1, OA_D);
      s1 = (pm >= pd) ? TS_M : TS_D;
      k--; break; }
    case TS_I: {                                           // select_i (:345-358)
      float pm = (TF(5, k) == 0.f) ? -INFINITY : OA(i - 1, k, OA_M);
      float pI = (TF(6, k) == 0.f) ? -INFINITY : OA(i - 1, k, OA_I);
      s1 = (pm >= pI) ? TS_M : TS_I;
      i--; break; }
    case TS_N: s1 = (i == 0) ? TS_S : TS_N; break;
    case TS_C: {                                           // select_c (:368-375)
      float p0 = !loopC ? -INFINITY : oax[(size_t)(i - 1) * 6 + 4] + ppx[(size_t)i * 6 + 4];
      float p1 = !moveE ? -INFINITY : oax[(size_t)i * 6 + 0];
      s1 = (p0 > p1) ? TS_C : TS_E;
      break; }
    case TS_J: {                                           // select_j (:378-386)
      float p0 = !loopJ ? -INFINITY : oax[(size_t)(i - 1) * 6 + 2] + ppx[(size_t)i * 6 + 2];
      float p1 = !loopE ? -INFINITY : oax[(size_t)i * 6 + 0];
      s1 = (p0 > p1) ? TS_J : TS_E;
      break; }
    case TS_E: {                                           // select_e (:390-409): striped scan, M with >=, D with >
      // scan position of node kk: q = (kk-1) % Q, r = (kk-1) / Q; M cells of stripe q precede its D cells.
      // Sequential semantics: the winner is the maximum value; among equal values a later M replaces anything,
      // a D replaces nothing.  So: best M by (value, latest position); D wins only when strictly greater than every M
      // scanned before it -- which needs D > best M overall or D before that M... handled by comparing positions.
      float bm = -INFINITY; int bmpos = -1, bmk = 0;
      float bd = -INFINITY; int bdpos = 0x7fffffff, bdk = 0;
      for (int kk = lane + 1; kk <= M; kk += 32) {
        const int q = (kk - 1) % Q, r = (kk - 1) / Q;
        const int pos = q * 2 * a.lanes_f32 + r;
        float vm = OA(i, kk, OA_M), vd = OA(i, kk, OA_D);
        if (vm > bm || (vm == bm && pos > bmpos)) { bm = vm; bmpos = pos; bmk = kk; }
        const int dpos = pos + a.lanes_f32;
        if (vd > bd || (vd == bd && dpos < bdpos)) { bd = vd; bdpos = dpos; bdk = kk; }
      }
#pragma unroll
      for (int dlt = 16; dlt >= 1; dlt >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, bm, dlt); int op = __shfl_xor_sync(0xffffffffu, bmpos, dlt); int ok = __shfl_xor_sync(0xffffffffu, bmk, dlt);
        if (om > bm || (om == bm && op > bmpos)) { bm = om; bmpos = op; bmk = ok; }
        float od = __shfl_xor_sync(0xffffffffu, bd, dlt); int odp = __shfl_xor_sync(0xffffffffu, bdpos, dlt); int odk = __shfl_xor_sync(0xffffffffu, bdk, dlt);
        if (od > bd || (od == bd && odp < bdpos)) { bd = od; bdpos = odp; bdk = odk; }
      }
      // pad cells of the last vectors hold -inf and are scanned too: with every real cell at -inf the last M pad wins (k > M)
      if (bd > bm) { s1 = TS_D; k = bdk; }
      else if (bm == -INFINITY) { bad = true; }
      else { s1 = TS_M; k = bmk; }
      break; }
    case TS_B: {                                           // select_b (:413-421)
      float p0 = !moveN ? -INFINITY : oax[(size_t)i * 6 + 1];
      float p1 = !moveJ ? -INFINITY : oax[(size_t)i * 6 + 2];
      s1 = (p0 > p1) ? TS_N : TS_J;
      break; }
    default: bad = true; break;
    }
    if (bad || i < 0 || k < 0 || n >= max_steps) { bad = true; break; }
    float postprob = 0.f;                                  // get_postprob (:264-280)
    if (s1 == TS_M)      postprob = PP(i, k, PPP_M);
    else if (s1 == TS_I) postprob = PP(i, k, PPP_I);
    else if (s1 == s0 && s1 == TS_N) postprob = ppx[(size_t)i * 6 + 1];
    else if (s1 == s0 && s1 == TS_C) postprob = ppx[(size_t)i * 6 + 4];
    else if (s1 == s0 && s1 == TS_J) postprob = ppx[(size_t)i * 6 + 2];
    emit(s1, k, i, postprob);
    if ((s1 == TS_N || s1 == TS_J || s1 == TS_C) && s1 == s0) i--;
    s0 = s1;
  }
  __syncwarp();
  if (bad) { if (lane == 0) { t.tlen[e] = 0; a.status[e] = 11; } return; }
  if (lane == 0) t.tlen[e] = n;
}

}  // namespace bathgpu
