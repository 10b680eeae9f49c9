/* fs_fwdback.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 *
 * Scalar, k-ordered restatement of the probability-space frameshift Forward and
 * Backward of src/impl_sse/fwdback_fs.c:
 *   p7_ForwardParser_Frameshift_3Codons   :97-533
 *   p7_BackwardParser_Frameshift_3Codons  :565-1013
 *   p7_Forward_Frameshift  (5 codons)     :2054-2607
 *   p7_Backward_Frameshift (5 codons)     :2634-2975
 * Semantics kept: ring depths, i-2 / i-1 look-backs, rescale trigger xE > 1e4 and
 * what gets rescaled, the un-rescaled x-buffers in the backward init rows, the
 * "committed scale" convention + insert_adj/adjN of the full matrices, the
 * has_own_scales switch, and the final-score formulas.  The D->D chain is the
 * plain serial recurrence the SIMD wrap passes converge to (:415-453).
 * Per-cell operation order follows the SIMD code; cross-k sums run in k order.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "bath_oracle.h"

#define TF(t,k) (om->tfv[(size_t)(t) * (M+1) + (k)])
#define RF(c,k) (om->rfv[(size_t)(c) * (M+1) + (k)])
#define XMX(mx,i,s) ((mx)->xmx[(size_t)(i) * BO_NXCELLS + (s)])

#define ROWS_FWD 4   /* PARSER_ROWS_FWD, hmmer.h:1149 */
#define ROWS_BWD 6   /* PARSER_ROWS_BWD, hmmer.h:1150 */

static inline int nuc3(uint8_t d) { return (d < BO_MAXNUC) ? d : BO_MAXCODONS3; }
static inline int nuc5(uint8_t d) { return (d < BO_MAXNUC) ? d : BO_MAXCODONS5; }
static inline int pmod(int a, int n) { return ((a % n) + n) % n; }

/* ------------------------------------------------------------------------ */
int bo_ForwardParser_Frameshift_3Codons(const uint8_t *dsq, int L, const BO_FS_OPROFILE *om, BO_MX *ox, float *opt_sc)
{
  int    M = om->M;
  float *mem, *mmx[ROWS_FWD], *imx[ROWS_FWD], *dmx[ROWS_FWD], *ivx[3];
  float  xN, xE, xB, xC, xJ;
  float  xN_buf[ROWS_FWD], xB_buf[ROWS_FWD], xJ_buf[ROWS_FWD], xC_buf[ROWS_FWD];
  int    i, k, r, u, v, w, x, c2, c3, c4;
  double totscale = 0.0;
  const float tNL = om->xf[BO_X_N][BO_O_LOOP], tNM = om->xf[BO_X_N][BO_O_MOVE];
  const float tJL = om->xf[BO_X_J][BO_O_LOOP], tJM = om->xf[BO_X_J][BO_O_MOVE];
  const float tCL = om->xf[BO_X_C][BO_O_LOOP], tCM = om->xf[BO_X_C][BO_O_MOVE];
  const float tEL = om->xf[BO_X_E][BO_O_LOOP], tEM = om->xf[BO_X_E][BO_O_MOVE];

  if (om->codon_lengths != 3) return BO_EINVAL;
  if (L < 3 || ox->allocL < L) return BO_EINVAL;

  mem = calloc((size_t)(3 * ROWS_FWD + 3) * (M + 2), sizeof(float));
  if (!mem) return BO_EMEM;
  for (r = 0; r < ROWS_FWD; r++) {
    mmx[r] = mem + (size_t)(3 * r + 0) * (M + 2);
    imx[r] = mem + (size_t)(3 * r + 1) * (M + 2);
    dmx[r] = mem + (size_t)(3 * r + 2) * (M + 2);
  }
  for (r = 0; r < 3; r++) ivx[r] = mem + (size_t)(3 * ROWS_FWD + r) * (M + 2);

  ox->M = M; ox->L = L; ox->has_own_scales = 1;

  for (r = 0; r < ROWS_FWD; r++) xN_buf[r] = xB_buf[r] = xJ_buf[r] = xC_buf[r] = 0.0f;
  xN_buf[0] = xN_buf[1] = 1.0f;
  xB_buf[0] = xB_buf[1] = tNM;
  for (i = 0; i <= 1; i++) {
    XMX(ox, i, BO_XC_SCALE) = 1.0f; XMX(ox, i, BO_XC_E) = 0.0f; XMX(ox, i, BO_XC_N) = 1.0f;
    XMX(ox, i, BO_XC_J) = 0.0f;     XMX(ox, i, BO_XC_B) = tNM;  XMX(ox, i, BO_XC_C) = 0.0f;
  }

  u = v = BO_MAXCODONS3;
  w = nuc3(dsq[1]);
  x = nuc3(dsq[2]);

  for (i = 2; i <= L; i++)
    {
      int curr  = i % ROWS_FWD;
      int prev2 = pmod(i - 2, ROWS_FWD);
      int prev3 = pmod(i - 3, ROWS_FWD);
      int ivx_2 = i % 3, ivx_3 = pmod(i - 1, 3), ivx_4 = pmod(i - 2, 3);
      float *mmc = mmx[curr], *imc = imx[curr], *dmc = dmx[curr];
      const float *mm2 = mmx[prev2], *im2 = imx[prev2], *dm2 = dmx[prev2];
      const float *mm3 = mmx[prev3], *im3 = imx[prev3];
      float xB2 = xB_buf[prev2];

      if (i > 2) { u = v; v = w; w = x; x = nuc3(dsq[i]); }
      c2 = BO_CODON2_FS3(w, x);       c2 = BO_MINIDX(c2, BO_DEGEN3_QC1);
      c3 = BO_CODON3_FS3(v, w, x);    c3 = BO_MINIDX(c3, BO_DEGEN3_C);
      c4 = BO_CODON4_FS3(u, v, w, x); c4 = BO_MINIDX(c4, BO_DEGEN3_QC1);

      xE = 0.0f;
      for (k = 1; k <= M; k++) {
        float sv, msv;
        sv  = xB2 * TF(BO_T_BM, k-1);
        sv  = sv + mm2[k-1] * TF(BO_T_MM, k-1);
        sv  = sv + im2[k-1] * TF(BO_T_IM, k-1);
        sv  = sv + dm2[k-1] * TF(BO_T_DM, k-1);
        ivx[ivx_2][k] = sv;
        msv = sv * RF(c2, k);
        if (i > 2) {   /* at i=2 the R3/R4 terms are not evaluated (:217-218) */
          msv = msv + ivx[ivx_3][k] * RF(c3, k);
          msv = msv + ivx[ivx_4][k] * RF(c4, k);
        }
        xE += msv;
        mmc[k] = msv;
        imc[k] = (i > 2) ? (mm3[k] * TF(BO_T_MI, k) + im3[k] * TF(BO_T_II, k)) : 0.0f;
      }
      dmc[1] = 0.0f;
      for (k = 2; k <= M; k++)
        dmc[k] = dmc[k-1] * TF(BO_T_DD, k-1) + mmc[k-1] * TF(BO_T_MD, k-1);
      for (k = 1; k <= M; k++) xE += dmc[k];

      if (i == 2) {
        xN = 1.0f;
        xJ = xE * tEL;
        xC = xE * tEM;
      } else {
        xN = xN_buf[prev3] * tNL;
        xJ = xJ_buf[prev3] * tJL + xE * tEL;
        xC = xC_buf[prev3] * tCL + xE * tEM;
      }
      xB = xN * tNM + xJ * tJM;

      if (xE > 1.0e4f) {
        float sf = 1.0f / xE;
        xN *= sf; xJ *= sf; xC *= sf; xB *= sf;
        for (r = 0; r < ROWS_FWD; r++)
          for (k = 1; k <= M; k++) { mmx[r][k] *= sf; dmx[r][k] *= sf; imx[r][k] *= sf; }
        for (r = 0; r < 3; r++)
          for (k = 1; k <= M; k++) ivx[r][k] *= sf;
        for (r = 0; r < ROWS_FWD; r++) { xN_buf[r] *= sf; xB_buf[r] *= sf; xJ_buf[r] *= sf; xC_buf[r] *= sf; }
        XMX(ox, i, BO_XC_SCALE) = xE;
        totscale += log(xE);
        xE = 1.0f;
      } else XMX(ox, i, BO_XC_SCALE) = 1.0f;

      xN_buf[curr] = xN; xB_buf[curr] = xB; xJ_buf[curr] = xJ; xC_buf[curr] = xC;
      XMX(ox, i, BO_XC_E) = xE; XMX(ox, i, BO_XC_N) = xN; XMX(ox, i, BO_XC_J) = xJ;
      XMX(ox, i, BO_XC_B) = xB; XMX(ox, i, BO_XC_C) = xC;
    }
  ox->totscale = (float) totscale;

  {
    float xCL   = xC_buf[L % ROWS_FWD];
    float xCLm1 = xC_buf[pmod(L - 1, ROWS_FWD)];
    float xCLm2 = xC_buf[pmod(L - 2, ROWS_FWD)];
    float xCtot = xCL + xCLm1 * tCL + xCLm2 * tCL;
    free(mem);
    if (isnan(xCtot) || isinf(xCtot)) return BO_ERANGE;
    if (L > 2 && xCtot == 0.0f) { if (opt_sc) *opt_sc = -INFINITY; return BO_ERANGE; }
    if (opt_sc) *opt_sc = ox->totscale + logf(xCtot * tCM);
  }
  return BO_OK;
}

/* ------------------------------------------------------------------------ */
/* one backward row's MDI given E, ivxf[1..M+1], I3 (may be NULL => zeros).
 * :859-909 */
static void bck_row_mdi(const BO_FS_OPROFILE *om, float xE, const float *ivxf, const float *i3, float adj3,
                        float *mmc, float *dmc, float *imc)
{
  int M = om->M, k;
  for (k = 1; k <= M; k++) {
    float ii = i3 ? i3[k] * adj3 : 0.0f;
    mmc[k] = xE   + ii * TF(BO_T_MI, k);
    imc[k] = 0.0f + ii * TF(BO_T_II, k);
    dmc[k] = xE;
  }
  if (ivxf)
    for (k = M; k >= 1; k--) {
      float carry = (k < M) ? ivxf[k+1] : 0.0f;
      mmc[k] = mmc[k] + carry * TF(BO_T_MM, k);
      imc[k] = imc[k] + carry * TF(BO_T_IM, k);
      dmc[k] = dmc[k] + carry * TF(BO_T_DM, k);
    }
  for (k = M - 1; k >= 1; k--)
    dmc[k] = dmc[k] + dmc[k+1] * TF(BO_T_DD, k);
  for (k = M - 1; k >= 1; k--)
    mmc[k] = mmc[k] + dmc[k+1] * TF(BO_T_MD, k);
}

int bo_BackwardParser_Frameshift_3Codons(const uint8_t *dsq, int L, const BO_FS_OPROFILE *om, const BO_MX *fwd, BO_MX *bck, float *opt_sc)
{
  int    M = om->M;
  float *mem, *mmx[ROWS_BWD], *imx[ROWS_BWD], *dmx[ROWS_BWD], *ivxf;
  float  xN, xE, xB, xC, xJ;
  float  xN_buf[ROWS_BWD], xB_buf[ROWS_BWD], xJ_buf[ROWS_BWD], xC_buf[ROWS_BWD];
  int    i, k, r, u, v, w, x, c2, c3, c4, b, b3;
  float  scale;
  double totscale = 0.0;
  const float tNL = om->xf[BO_X_N][BO_O_LOOP], tNM = om->xf[BO_X_N][BO_O_MOVE];
  const float tJL = om->xf[BO_X_J][BO_O_LOOP], tJM = om->xf[BO_X_J][BO_O_MOVE];
  const float tCL = om->xf[BO_X_C][BO_O_LOOP], tCM = om->xf[BO_X_C][BO_O_MOVE];
  const float tEL = om->xf[BO_X_E][BO_O_LOOP], tEM = om->xf[BO_X_E][BO_O_MOVE];

  if (om->codon_lengths != 3) return BO_EINVAL;
  if (L < 5 || bck->allocL < L) return BO_EINVAL;

  mem = calloc((size_t)(3 * ROWS_BWD + 1) * (M + 2), sizeof(float));
  if (!mem) return BO_EMEM;
  for (r = 0; r < ROWS_BWD; r++) {
    mmx[r] = mem + (size_t)(3 * r + 0) * (M + 2);
    imx[r] = mem + (size_t)(3 * r + 1) * (M + 2);
    dmx[r] = mem + (size_t)(3 * r + 2) * (M + 2);
  }
  ivxf = mem + (size_t)(3 * ROWS_BWD) * (M + 2);

  bck->M = M; bck->L = L; bck->has_own_scales = 0;
  for (r = 0; r < ROWS_BWD; r++) xN_buf[r] = xB_buf[r] = xJ_buf[r] = xC_buf[r] = 0.0f;

  /* rows L and L-1 (:628-690) */
  for (i = L; i >= L - 1; i--) {
    b = i % ROWS_BWD;
    xC = (i == L) ? tCM : tCL * tCM;
    xN = xB = xJ = 0.0f;
    xE = xC * tEM;
    bck_row_mdi(om, xE, NULL, NULL, 1.0f, mmx[b], dmx[b], imx[b]);
    scale = XMX(fwd, i, BO_XC_SCALE);
    XMX(bck, i, BO_XC_SCALE) = scale;
    if (scale > 1.0f) {
      float sf = 1.0f / scale;
      xN *= sf; xJ *= sf; xC *= sf; xB *= sf; xE *= sf;
      for (r = 0; r < ROWS_BWD; r++)
        for (k = 1; k <= M; k++) { mmx[r][k] *= sf; dmx[r][k] *= sf; imx[r][k] *= sf; }
      /* x-buffers are NOT rescaled in these two rows (:673-678) */
      totscale += log(scale);
    }
    xN_buf[b] = xN; xB_buf[b] = xB; xJ_buf[b] = xJ; xC_buf[b] = xC;
    XMX(bck, i, BO_XC_E) = xE; XMX(bck, i, BO_XC_N) = xN; XMX(bck, i, BO_XC_J) = xJ;
    XMX(bck, i, BO_XC_B) = xB; XMX(bck, i, BO_XC_C) = xC;
  }

  u = v = BO_MAXCODONS3;
  w = nuc3(dsq[L]);
  x = nuc3(dsq[L-1]);

  for (i = L - 2; i >= 1; i--)
    {
      int prev2 = (i + 2) % ROWS_BWD, prev3 = (i + 3) % ROWS_BWD, prev4 = (i + 4) % ROWS_BWD;
      b = i % ROWS_BWD; b3 = (i + 3) % ROWS_BWD;

      if (i < L - 2) { u = v; v = w; w = x; x = nuc3(dsq[i+1]); }
      c2 = BO_CODON2_FS3(x, w);       c2 = BO_MINIDX(c2, BO_DEGEN3_QC1);
      c3 = BO_CODON3_FS3(x, w, v);    c3 = BO_MINIDX(c3, BO_DEGEN3_C);
      c4 = BO_CODON4_FS3(x, w, v, u); c4 = BO_MINIDX(c4, BO_DEGEN3_QC1);

      xB = 0.0f;
      if (i == L - 2) {
        for (k = 1; k <= M; k++) ivxf[k] = mmx[prev2][k] * RF(c2, k);
      } else {
        for (k = 1; k <= M; k++)
          ivxf[k] = (mmx[prev2][k] * RF(c2, k) + mmx[prev3][k] * RF(c3, k)) + mmx[prev4][k] * RF(c4, k);
      }
      for (k = 1; k <= M; k++) xB += ivxf[k] * TF(BO_T_BM, k-1);

      if (i == L - 2) {
        xC = tCL * tCM;
        xJ = xB * tJM;
        xN = xB * tNM;
      } else {
        xC = xC_buf[b3] * tCL;
        xJ = xJ_buf[b3] * tJL + xB * tJM;
        xN = xN_buf[b3] * tNL + xB * tNM;
      }
      xE = xJ * tEL + xC * tEM;

      bck_row_mdi(om, xE, ivxf, (i == L - 2) ? NULL : imx[prev3], 1.0f, mmx[b], dmx[b], imx[b]);

      if (i == L - 2) scale = XMX(fwd, i, BO_XC_SCALE);
      else {
        if (xB > 1.0e16f) bck->has_own_scales = 1;
        if (bck->has_own_scales) scale = (xB > 1.0e4f) ? xB : 1.0f;
        else                     scale = XMX(fwd, i, BO_XC_SCALE);
      }
      XMX(bck, i, BO_XC_SCALE) = scale;
      if (scale > 1.0f) {
        float sf = 1.0f / scale;
        xN *= sf; xJ *= sf; xC *= sf; xB *= sf; xE *= sf;
        for (r = 0; r < ROWS_BWD; r++)
          for (k = 1; k <= M; k++) { mmx[r][k] *= sf; dmx[r][k] *= sf; imx[r][k] *= sf; }
        for (r = 0; r < ROWS_BWD; r++) { xN_buf[r] *= sf; xB_buf[r] *= sf; xJ_buf[r] *= sf; xC_buf[r] *= sf; }
        totscale += log(scale);
      }
      xN_buf[b] = xN; xB_buf[b] = xB; xJ_buf[b] = xJ; xC_buf[b] = xC;
      XMX(bck, i, BO_XC_E) = xE; XMX(bck, i, BO_XC_N) = xN; XMX(bck, i, BO_XC_J) = xJ;
      XMX(bck, i, BO_XC_B) = xB; XMX(bck, i, BO_XC_C) = xC;
    }

  /* termination, i = 0 (:951-987) */
  u = v; v = w; w = x; x = nuc3(dsq[1]);
  c2 = BO_CODON2_FS3(x, w);       c2 = BO_MINIDX(c2, BO_DEGEN3_QC1);
  c3 = BO_CODON3_FS3(x, w, v);    c3 = BO_MINIDX(c3, BO_DEGEN3_C);
  c4 = BO_CODON4_FS3(x, w, v, u); c4 = BO_MINIDX(c4, BO_DEGEN3_QC1);
  xB = 0.0f;
  for (k = 1; k <= M; k++) {
    ivxf[k] = (mmx[2][k] * RF(c2, k) + mmx[3][k] * RF(c3, k)) + mmx[4][k] * RF(c4, k);
    xB += ivxf[k] * TF(BO_T_BM, k-1);
  }
  xN = xN_buf[3] * tNL + xB * tNM;
  XMX(bck, 0, BO_XC_B) = xB; XMX(bck, 0, BO_XC_N) = xN; XMX(bck, 0, BO_XC_J) = 0.0f;
  XMX(bck, 0, BO_XC_C) = 0.0f; XMX(bck, 0, BO_XC_E) = 0.0f; XMX(bck, 0, BO_XC_SCALE) = 1.0f;
  bck->totscale = (float) totscale;

  {
    float xNtot = xN + xN_buf[1] + xN_buf[2];
    free(mem);
    if (isnan(xNtot) || isinf(xNtot)) return BO_ERANGE;
    if (L > 0 && xNtot == 0.0f) { if (opt_sc) *opt_sc = -INFINITY; return BO_ERANGE; }
    if (opt_sc) *opt_sc = bck->totscale + logf(xNtot);
  }
  return BO_OK;
}

/* ------------------------------------------------------------------------ */
#define FCELL(mx,i,k,s) ((mx)->dp[((size_t)(i) * (M+1) + (k)) * BO_NSCELLS_FS + (s)])
#define BCELL(mx,i,k,s) ((mx)->dp[((size_t)(i) * (M+1) + (k)) * BO_NSCELLS    + (s)])

int bo_Forward_Frameshift(const uint8_t *dsq, int L, const BO_FS_OPROFILE *om, BO_MX *fwd, float *opt_sc)
{
  int    M = om->M;
  float *mem, *ivx[5];
  float  xN, xE, xB, xC, xJ;
  float  xN_buf[ROWS_FWD], xB_buf[ROWS_FWD], xJ_buf[ROWS_FWD], xC_buf[ROWS_FWD];
  int    i, k, r, t, u, v, w, x, c1, c2, c3, c4, c5;
  double totscale = 0.0;
  const float tNL = om->xf[BO_X_N][BO_O_LOOP], tNM = om->xf[BO_X_N][BO_O_MOVE];
  const float tJL = om->xf[BO_X_J][BO_O_LOOP], tJM = om->xf[BO_X_J][BO_O_MOVE];
  const float tCL = om->xf[BO_X_C][BO_O_LOOP], tCM = om->xf[BO_X_C][BO_O_MOVE];
  const float tEL = om->xf[BO_X_E][BO_O_LOOP], tEM = om->xf[BO_X_E][BO_O_MOVE];

  if (om->codon_lengths != 5) return BO_EINVAL;
  if (fwd->nscells != BO_NSCELLS_FS || fwd->allocL < L || fwd->M != M || L < 2) return BO_EINVAL;

  mem = calloc((size_t) 5 * (M + 2), sizeof(float));
  if (!mem) return BO_EMEM;
  for (r = 0; r < 5; r++) ivx[r] = mem + (size_t) r * (M + 2);

  fwd->L = L; fwd->has_own_scales = 1;
  memset(&FCELL(fwd, 0, 0, 0), 0, sizeof(float) * (size_t)(M + 1) * BO_NSCELLS_FS);

  for (r = 0; r < ROWS_FWD; r++) xN_buf[r] = xB_buf[r] = xJ_buf[r] = xC_buf[r] = 0.0f;
  xN_buf[0] = xN_buf[1] = xN_buf[2] = 1.0f;
  xB_buf[0] = xB_buf[1] = xB_buf[2] = tNM;
  for (r = 0; r < 3 && r <= L; r++) {
    XMX(fwd, r, BO_XC_SCALE) = 1.0f; XMX(fwd, r, BO_XC_E) = 0.0f; XMX(fwd, r, BO_XC_N) = 1.0f;
    XMX(fwd, r, BO_XC_J) = 0.0f;     XMX(fwd, r, BO_XC_B) = tNM;  XMX(fwd, r, BO_XC_C) = 0.0f;
  }

  t = u = v = w = x = BO_MAXCODONS5;
  for (i = 1; i <= L; i++)
    {
      int ivx_1 = i % 5, ivx_2 = pmod(i-1, 5), ivx_3 = pmod(i-2, 5), ivx_4 = pmod(i-3, 5), ivx_5 = pmod(i-4, 5);
      int b = i % ROWS_FWD, b1 = pmod(i-1, ROWS_FWD), b3 = pmod(i-3, ROWS_FWD);
      float xB1 = xB_buf[b1];
      float insert_adj = 1.0f;

      if (i <= 2) { t = u = v = BO_MAXCODONS5; w = (i == 2) ? x : BO_MAXCODONS5; x = nuc5(dsq[i]); }
      else        { t = u; u = v; v = w; w = x; x = nuc5(dsq[i]); }
      c1 = BO_CODON1_FS5(x);             c1 = BO_MINIDX(c1, BO_DEGEN5_QC2);
      c2 = BO_CODON2_FS5(w, x);          c2 = BO_MINIDX(c2, BO_DEGEN5_QC1);
      c3 = BO_CODON3_FS5(v, w, x);       c3 = BO_MINIDX(c3, BO_DEGEN5_C);
      c4 = BO_CODON4_FS5(u, v, w, x);    c4 = BO_MINIDX(c4, BO_DEGEN5_QC1);
      c5 = BO_CODON5_FS5(t, u, v, w, x); c5 = BO_MINIDX(c5, BO_DEGEN5_QC2);

      if (i >= 3) insert_adj = 1.0f / (XMX(fwd, i-2, BO_XC_SCALE) * XMX(fwd, i-1, BO_XC_SCALE));

      xE = 0.0f;
      FCELL(fwd, i, 0, BO_FS_D) = FCELL(fwd, i, 0, BO_FS_I) = 0.0f;
      for (r = 0; r < 6; r++) FCELL(fwd, i, 0, BO_FS_M + r) = 0.0f;
      for (k = 1; k <= M; k++) {
        float sv, mc1, mc2 = 0.0f, mc3 = 0.0f, mc4 = 0.0f, mc5 = 0.0f, msv;
        sv = xB1 * TF(BO_T_BM, k-1);
        sv = sv + FCELL(fwd, i-1, k-1, BO_FS_M) * TF(BO_T_MM, k-1);
        sv = sv + FCELL(fwd, i-1, k-1, BO_FS_I) * TF(BO_T_IM, k-1);
        sv = sv + FCELL(fwd, i-1, k-1, BO_FS_D) * TF(BO_T_DM, k-1);
        ivx[ivx_1][k] = sv;
        mc1 = sv * RF(c1, k);
        if (i == 1)      msv = mc1;
        else if (i == 2) { mc2 = ivx[ivx_2][k] * RF(c2, k); msv = mc1 + mc2; }
        else {
          mc2 = ivx[ivx_2][k] * RF(c2, k);
          mc3 = ivx[ivx_3][k] * RF(c3, k);
          mc4 = ivx[ivx_4][k] * RF(c4, k);
          mc5 = ivx[ivx_5][k] * RF(c5, k);
          msv = ((mc1 + mc2) + (mc3 + mc4)) + mc5;
        }
        xE += msv;
        FCELL(fwd, i, k, BO_FS_M + 0) = msv;
        FCELL(fwd, i, k, BO_FS_M + 1) = mc1;
        FCELL(fwd, i, k, BO_FS_M + 2) = mc2;
        FCELL(fwd, i, k, BO_FS_M + 3) = mc3;
        FCELL(fwd, i, k, BO_FS_M + 4) = mc4;
        FCELL(fwd, i, k, BO_FS_M + 5) = mc5;
        if (i >= 3)
          FCELL(fwd, i, k, BO_FS_I) = (FCELL(fwd, i-3, k, BO_FS_M) * insert_adj) * TF(BO_T_MI, k)
                                    + (FCELL(fwd, i-3, k, BO_FS_I) * insert_adj) * TF(BO_T_II, k);
        else
          FCELL(fwd, i, k, BO_FS_I) = 0.0f;
      }
      FCELL(fwd, i, 1, BO_FS_D) = 0.0f;
      for (k = 2; k <= M; k++)
        FCELL(fwd, i, k, BO_FS_D) = FCELL(fwd, i, k-1, BO_FS_D) * TF(BO_T_DD, k-1) + FCELL(fwd, i, k-1, BO_FS_M) * TF(BO_T_MD, k-1);
      for (k = 1; k <= M; k++) xE += FCELL(fwd, i, k, BO_FS_D);

      if (i <= 2) {
        xN = 1.0f;
        xJ = xE * tEL;
        xC = xE * tEM;
      } else {
        xN = xN_buf[b3] * tNL;
        xJ = xJ_buf[b3] * tJL + xE * tEL;
        xC = xC_buf[b3] * tCL + xE * tEM;
      }
      xB = xN * tNM + xJ * tJM;

      if (xE > 1.0e4f) {
        float sf = 1.0f / xE;
        xN *= sf; xJ *= sf; xC *= sf; xB *= sf;
        for (k = 1; k <= M; k++)
          for (r = 0; r < BO_NSCELLS_FS; r++) FCELL(fwd, i, k, r) *= sf;
        for (r = 0; r < 5; r++)
          for (k = 1; k <= M; k++) ivx[r][k] *= sf;
        for (r = 0; r < ROWS_FWD; r++) { xN_buf[r] *= sf; xB_buf[r] *= sf; xJ_buf[r] *= sf; xC_buf[r] *= sf; }
        XMX(fwd, i, BO_XC_SCALE) = xE;
        totscale += log(xE);
        xE = 1.0f;
      } else XMX(fwd, i, BO_XC_SCALE) = 1.0f;

      xN_buf[b] = xN; xB_buf[b] = xB; xJ_buf[b] = xJ; xC_buf[b] = xC;
      XMX(fwd, i, BO_XC_E) = xE; XMX(fwd, i, BO_XC_N) = xN; XMX(fwd, i, BO_XC_J) = xJ;
      XMX(fwd, i, BO_XC_B) = xB; XMX(fwd, i, BO_XC_C) = xC;
    }
  fwd->totscale = (float) totscale;
  free(mem);

  {
    float xCL   = xC_buf[L % ROWS_FWD];
    float xCLm1 = xC_buf[pmod(L - 1, ROWS_FWD)];
    float xCLm2 = xC_buf[pmod(L - 2, ROWS_FWD)];
    float xCtot = xCL + xCLm1 * tCL + xCLm2 * tCL;
    if (isnan(xCtot) || isinf(xCtot)) return BO_ERANGE;
    if (L > 1 && xCtot == 0.0f) { if (opt_sc) *opt_sc = -INFINITY; return BO_ERANGE; }
    if (opt_sc) *opt_sc = fwd->totscale + logf(xCtot * tCM);
  }
  return BO_OK;
}

/* ------------------------------------------------------------------------ */
int bo_Backward_Frameshift(const uint8_t *dsq, int L, const BO_FS_OPROFILE *om, const BO_MX *fwd, BO_MX *bck, float *opt_sc)
{
  int    M = om->M;
  float *mem, *ivxf, *mmc, *dmc, *imc, *i3row;
  float  xN, xE, xB, xC, xJ;
  float  xN_buf[ROWS_BWD], xB_buf[ROWS_BWD], xJ_buf[ROWS_BWD], xC_buf[ROWS_BWD];
  int    i, k, r, t, u, v, w, x, c1, c2, c3, c4, c5, b, b3;
  float  scale, adj2, adj3, adj4, adj5;
  double totscale = 0.0;
  const float tNL = om->xf[BO_X_N][BO_O_LOOP], tNM = om->xf[BO_X_N][BO_O_MOVE];
  const float tJL = om->xf[BO_X_J][BO_O_LOOP], tJM = om->xf[BO_X_J][BO_O_MOVE];
  const float tCL = om->xf[BO_X_C][BO_O_LOOP], tCM = om->xf[BO_X_C][BO_O_MOVE];
  const float tEL = om->xf[BO_X_E][BO_O_LOOP], tEM = om->xf[BO_X_E][BO_O_MOVE];

  if (om->codon_lengths != 5) return BO_EINVAL;
  if (bck->nscells != BO_NSCELLS || bck->allocL < L || bck->M != M || L < 2) return BO_EINVAL;

  mem = calloc((size_t) 5 * (M + 2), sizeof(float));
  if (!mem) return BO_EMEM;
  ivxf = mem; mmc = mem + (M + 2); dmc = mem + 2 * (M + 2); imc = mem + 3 * (M + 2); i3row = mem + 4 * (M + 2);

  bck->L = L; bck->has_own_scales = 0;
  for (r = 0; r < ROWS_BWD; r++) xN_buf[r] = xB_buf[r] = xJ_buf[r] = xC_buf[r] = 0.0f;
  xC_buf[(L + 1) % ROWS_BWD] = tCM;
  xC_buf[(L + 2) % ROWS_BWD] = tCM;

  /* row L (:2689-2741) */
  i = L; b = i % ROWS_BWD;
  xC = tCM; xN = xB = xJ = 0.0f;
  xE = xC * tEM;
  bck_row_mdi(om, xE, NULL, NULL, 1.0f, mmc, dmc, imc);
  scale = XMX(fwd, L, BO_XC_SCALE);
  XMX(bck, L, BO_XC_SCALE) = scale;
  if (scale > 1.0f) {
    float sf = 1.0f / scale;
    xN *= sf; xJ *= sf; xC *= sf; xB *= sf; xE *= sf;
    for (k = 1; k <= M; k++) { mmc[k] *= sf; dmc[k] *= sf; imc[k] *= sf; }
    totscale += log(scale);
  }
  for (k = 0; k <= M; k++) {
    BCELL(bck, L, k, BO_S_M) = (k ? mmc[k] : 0.0f);
    BCELL(bck, L, k, BO_S_D) = (k ? dmc[k] : 0.0f);
    BCELL(bck, L, k, BO_S_I) = (k ? imc[k] : 0.0f);
  }
  xN_buf[b] = xN; xB_buf[b] = xB; xJ_buf[b] = xJ; xC_buf[b] = xC;
  XMX(bck, L, BO_XC_E) = xE; XMX(bck, L, BO_XC_N) = xN; XMX(bck, L, BO_XC_J) = xJ;
  XMX(bck, L, BO_XC_B) = xB; XMX(bck, L, BO_XC_C) = xC;

  t = u = v = w = x = BO_MAXCODONS5;
  for (i = L - 1; i >= 0; i--)
    {
      if (i >= 1) {
        t = u; u = v; v = w; w = x; x = nuc5(dsq[i+1]);
      } else {   /* termination reloads the window explicitly (:2893-2897) */
        x = nuc5(dsq[1]);
        w = (L >= 2) ? nuc5(dsq[2]) : BO_MAXCODONS5;
        v = (L >= 3) ? nuc5(dsq[3]) : BO_MAXCODONS5;
        u = (L >= 4) ? nuc5(dsq[4]) : BO_MAXCODONS5;
        t = (L >= 5) ? nuc5(dsq[5]) : BO_MAXCODONS5;
      }
      c1 = BO_CODON1_FS5(x);             c1 = BO_MINIDX(c1, BO_DEGEN5_QC2);
      c2 = BO_CODON2_FS5(x, w);          c2 = BO_MINIDX(c2, BO_DEGEN5_QC1);
      c3 = BO_CODON3_FS5(x, w, v);       c3 = BO_MINIDX(c3, BO_DEGEN5_C);
      c4 = BO_CODON4_FS5(x, w, v, u);    c4 = BO_MINIDX(c4, BO_DEGEN5_QC1);
      c5 = BO_CODON5_FS5(x, w, v, u, t); c5 = BO_MINIDX(c5, BO_DEGEN5_QC2);

      adj2 = (i + 2 <= L) ? 1.0f / XMX(fwd, i+1, BO_XC_SCALE) : 1.0f;
      adj3 = (i + 3 <= L) ? adj2 / XMX(fwd, i+2, BO_XC_SCALE) : 1.0f;
      adj4 = (i + 4 <= L) ? adj3 / XMX(fwd, i+3, BO_XC_SCALE) : 1.0f;
      adj5 = (i + 5 <= L) ? adj4 / XMX(fwd, i+4, BO_XC_SCALE) : 1.0f;

      xB = 0.0f;
      for (k = 1; k <= M; k++) {
        float m1 = BCELL(bck, i+1, k, BO_S_M);
        float m2 = (i + 2 <= L) ? BCELL(bck, i+2, k, BO_S_M) : 0.0f;
        float m3 = (i + 3 <= L) ? BCELL(bck, i+3, k, BO_S_M) : 0.0f;
        float m4 = (i + 4 <= L) ? BCELL(bck, i+4, k, BO_S_M) : 0.0f;
        float m5 = (i + 5 <= L) ? BCELL(bck, i+5, k, BO_S_M) : 0.0f;
        ivxf[k] = ((m1 * RF(c1, k) + (m2 * adj2) * RF(c2, k)) +
                   ((m3 * adj3) * RF(c3, k) + (m4 * adj4) * RF(c4, k))) +
                  (m5 * adj5) * RF(c5, k);
        xB += ivxf[k] * TF(BO_T_BM, k-1);
      }

      if (i == 0) {
        xN = xN_buf[3] * tNL + xB * tNM;
        XMX(bck, 0, BO_XC_B) = xB; XMX(bck, 0, BO_XC_N) = xN; XMX(bck, 0, BO_XC_J) = 0.0f;
        XMX(bck, 0, BO_XC_C) = 0.0f; XMX(bck, 0, BO_XC_E) = 0.0f; XMX(bck, 0, BO_XC_SCALE) = 1.0f;
        memset(&BCELL(bck, 0, 0, 0), 0, sizeof(float) * (size_t)(M + 1) * BO_NSCELLS);
        break;
      }

      b = i % ROWS_BWD; b3 = (i + 3) % ROWS_BWD;
      xC = xC_buf[b3] * tCL;
      xJ = xJ_buf[b3] * tJL + xB * tJM;
      xN = xN_buf[b3] * tNL + xB * tNM;
      xE = xJ * tEL + xC * tEM;

      if (i + 3 <= L) { for (k = 1; k <= M; k++) i3row[k] = BCELL(bck, i+3, k, BO_S_I); }
      else            { for (k = 1; k <= M; k++) i3row[k] = 0.0f; }
      bck_row_mdi(om, xE, ivxf, i3row, adj3, mmc, dmc, imc);

      if (bck->has_own_scales) scale = (xB > 1.0e4f) ? xB : 1.0f;
      else                     scale = XMX(fwd, i, BO_XC_SCALE);
      if (xB > 1.0e16f)        bck->has_own_scales = 1;

      XMX(bck, i, BO_XC_SCALE) = scale;
      if (scale > 1.0f) {
        float sf = 1.0f / scale;
        xN *= sf; xJ *= sf; xC *= sf; xB *= sf; xE *= sf;
        for (k = 1; k <= M; k++) { mmc[k] *= sf; dmc[k] *= sf; imc[k] *= sf; }
        for (r = 0; r < ROWS_BWD; r++) { xN_buf[r] *= sf; xB_buf[r] *= sf; xJ_buf[r] *= sf; xC_buf[r] *= sf; }
        totscale += log(scale);
      }
      for (k = 0; k <= M; k++) {
        BCELL(bck, i, k, BO_S_M) = (k ? mmc[k] : 0.0f);
        BCELL(bck, i, k, BO_S_D) = (k ? dmc[k] : 0.0f);
        BCELL(bck, i, k, BO_S_I) = (k ? imc[k] : 0.0f);
      }
      xN_buf[b] = xN; xB_buf[b] = xB; xJ_buf[b] = xJ; xC_buf[b] = xC;
      XMX(bck, i, BO_XC_E) = xE; XMX(bck, i, BO_XC_N) = xN; XMX(bck, i, BO_XC_J) = xJ;
      XMX(bck, i, BO_XC_B) = xB; XMX(bck, i, BO_XC_C) = xC;
    }
  bck->totscale = (float) totscale;
  free(mem);

  {
    float xNtot = xN + xN_buf[1] + xN_buf[2];
    if (isnan(xNtot) || isinf(xNtot)) return BO_ERANGE;
    if (L > 0 && xNtot == 0.0f) { if (opt_sc) *opt_sc = -INFINITY; return BO_ERANGE; }
    if (opt_sc) *opt_sc = bck->totscale + logf(xNtot);
  }
  return BO_OK;
}
