/* orf_domain.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 * The standard-translation branch's DP over an ORF (amino acid sequence), as production runs it in src/impl_sse:
 * probability space, sparse rescaling, restated un-striped in k order.
 *   bo_Forward / bo_Backward          forward_engine / backward_engine  (src/impl_sse/fwdback.c:256-466, :468-738)
 *   bo_Decoding / bo_DomainDecoding   src/impl_sse/decoding.c:76-139, :160-196
 *   bo_OptimalAccuracy / bo_OATrace   src/impl_sse/optacc.c:58-174, :225-425
 *   bo_Null2_ByExpectation            src/impl_sse/null2.c:44-125
 * Matrices are BO_MX with 3 cells per node (M, D, I); a parser call passes nscells == 0 and keeps the X rows only.
 * The D->D chains are the plain serial recurrences the SIMD passes converge to (fwdback.c:352-395 stops early
 * once a pass adds less than machine epsilon; the difference is below float resolution of the sums). */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "bath_oracle.h"

#define TF(t,k)      (om->tfv[(size_t)(t) * (M+1) + (k)])
#define XMX(mx,i,s)  ((mx)->xmx[(size_t)(i) * BO_NXCELLS + (s)])
#define CELL(mx,i,k,s) ((mx)->dp[((size_t)(i) * ((mx)->M + 1) + (k)) * BO_NSCELLS + (s)])

/* p7_oprofile_ReconfigMultihit / Unihit (src/impl_sse/p7_oprofile.c:1384-1429) */
void bo_oprofile_ReconfigMultihit(BO_OPROFILE *om, int L)
{
  om->xf[BO_X_E][BO_O_MOVE] = 0.5f; om->xf[BO_X_E][BO_O_LOOP] = 0.5f; om->nj = 1.0f;
  bo_oprofile_ReconfigLength(om, L);
}
void bo_oprofile_ReconfigUnihit(BO_OPROFILE *om, int L)
{
  om->xf[BO_X_E][BO_O_MOVE] = 1.0f; om->xf[BO_X_E][BO_O_LOOP] = 0.0f; om->nj = 0.0f;
  bo_oprofile_ReconfigLength(om, L);
}

/* forward_engine (fwdback.c:256-466).  ox->nscells == 3 keeps every row (p7_Forward), 0 keeps X rows only (p7_ForwardParser). */
int bo_Forward(const uint8_t *dsq, int L, const BO_OPROFILE *om, BO_MX *ox, float *opt_sc)
{
  const int full = ox->nscells == BO_NSCELLS;
  int    M = om->M, i, k;
  float *mp, *ip, *dp, *mc, *ic, *dc, *tmp;
  float  xN, xE, xB, xC, xJ;
  float *mem = calloc((size_t) 6 * (M + 2), sizeof(float));
  if (!mem) return BO_EMEM;
  mp = mem; ip = mem + (M + 2); dp = mem + 2 * (M + 2); mc = mem + 3 * (M + 2); ic = mem + 4 * (M + 2); dc = mem + 5 * (M + 2);

  ox->M = M; ox->L = L; ox->has_own_scales = 1;
  if (full) for (k = 0; k <= M; k++) CELL(ox, 0, k, BO_S_M) = CELL(ox, 0, k, BO_S_D) = CELL(ox, 0, k, BO_S_I) = 0.0f;
  xE = XMX(ox, 0, BO_XC_E) = 0.;
  xN = XMX(ox, 0, BO_XC_N) = 1.;
  xJ = XMX(ox, 0, BO_XC_J) = 0.;
  xB = XMX(ox, 0, BO_XC_B) = om->xf[BO_X_N][BO_O_MOVE];
  xC = XMX(ox, 0, BO_XC_C) = 0.;
  XMX(ox, 0, BO_XC_SCALE) = 1.0f;
  ox->totscale = 0.0f;

  for (i = 1; i <= L; i++) {
    const float *rf = om->rfv + (size_t) dsq[i] * (M + 1);
    xE = 0.0f;
    mc[0] = ic[0] = dc[0] = 0.0f;
    for (k = 1; k <= M; k++) {
      float sv = xB * TF(BO_T_BM, k-1);
      sv = sv + mp[k-1] * TF(BO_T_MM, k-1);
      sv = sv + ip[k-1] * TF(BO_T_IM, k-1);
      sv = sv + dp[k-1] * TF(BO_T_DM, k-1);
      sv = sv * rf[k];
      xE += sv;
      mc[k] = sv;
      ic[k] = mp[k] * TF(BO_T_MI, k) + ip[k] * TF(BO_T_II, k);
    }
    dc[1] = 0.0f;
    for (k = 2; k <= M; k++) dc[k] = mc[k-1] * TF(BO_T_MD, k-1) + dc[k-1] * TF(BO_T_DD, k-1);
    for (k = 1; k <= M; k++) xE += dc[k];

    xN = xN * om->xf[BO_X_N][BO_O_LOOP];
    xC = (xC * om->xf[BO_X_C][BO_O_LOOP]) + (xE * om->xf[BO_X_E][BO_O_MOVE]);
    xJ = (xJ * om->xf[BO_X_J][BO_O_LOOP]) + (xE * om->xf[BO_X_E][BO_O_LOOP]);
    xB = (xJ * om->xf[BO_X_J][BO_O_MOVE]) + (xN * om->xf[BO_X_N][BO_O_MOVE]);

    if (xE > 1.0e4) {
      float sf = 1.0 / xE;
      xN = xN / xE; xC = xC / xE; xJ = xJ / xE; xB = xB / xE;
      for (k = 1; k <= M; k++) { mc[k] *= sf; dc[k] *= sf; ic[k] *= sf; }
      XMX(ox, i, BO_XC_SCALE) = xE;
      ox->totscale += log(xE);
      xE = 1.0;
    } else XMX(ox, i, BO_XC_SCALE) = 1.0f;
    XMX(ox, i, BO_XC_E) = xE; XMX(ox, i, BO_XC_N) = xN; XMX(ox, i, BO_XC_J) = xJ;
    XMX(ox, i, BO_XC_B) = xB; XMX(ox, i, BO_XC_C) = xC;
    if (full) for (k = 0; k <= M; k++) { CELL(ox, i, k, BO_S_M) = mc[k]; CELL(ox, i, k, BO_S_D) = dc[k]; CELL(ox, i, k, BO_S_I) = ic[k]; }
    tmp = mp; mp = mc; mc = tmp;
    tmp = ip; ip = ic; ic = tmp;
    tmp = dp; dp = dc; dc = tmp;
  }
  free(mem);
  if (isnan(xC))             return BO_ERANGE;
  if (L > 0 && xC == 0.0)    { if (opt_sc) *opt_sc = -INFINITY; return BO_ERANGE; }
  if (isinf(xC))             return BO_ERANGE;
  if (opt_sc) *opt_sc = ox->totscale + log(xC * om->xf[BO_X_C][BO_O_MOVE]);
  return BO_OK;
}

/* backward_engine (fwdback.c:468-738) */
int bo_Backward(const uint8_t *dsq, int L, const BO_OPROFILE *om, const BO_MX *fwd, BO_MX *bck, float *opt_sc)
{
  const int full = bck->nscells == BO_NSCELLS;
  int    M = om->M, i, k;
  float  xN, xE, xB, xC, xJ, sc;
  float *mem = calloc((size_t) 6 * (M + 3), sizeof(float));
  float *mp, *ip, *dpv, *mc, *ic, *dc, *tmp;     /* "previous" = row i+1 */
  if (!mem) return BO_EMEM;
  mp = mem; ip = mem + (M + 3); dpv = mem + 2 * (M + 3); mc = mem + 3 * (M + 3); ic = mem + 4 * (M + 3); dc = mem + 5 * (M + 3);

  bck->M = M; bck->L = L; bck->has_own_scales = 0;
  xJ = 0.0; xB = 0.0; xN = 0.0;
  xC = om->xf[BO_X_C][BO_O_MOVE];
  xE = xC * om->xf[BO_X_E][BO_O_MOVE];
  /* row L (:491-529): D(L,k) = xE + D(L,k+1) tDD(k); M(L,k) = xE + D(L,k+1) tMD(k); I = 0 */
  mc[M+1] = dc[M+1] = ic[M+1] = 0.0f;
  for (k = M; k >= 1; k--) {
    dc[k] = xE + ((k < M) ? dc[k+1] * TF(BO_T_DD, k) : 0.0f);
    mc[k] = xE + ((k < M) ? dc[k+1] * TF(BO_T_MD, k) : 0.0f);
    ic[k] = 0.0f;
  }
  sc = XMX(fwd, L, BO_XC_SCALE);
  if (sc > 1.0f) {
    float sf = 1.0 / sc;
    xE = xE / sc; xN = xN / sc; xC = xC / sc; xJ = xJ / sc; xB = xB / sc;
    for (k = 1; k <= M; k++) { mc[k] *= sf; dc[k] *= sf; ic[k] *= sf; }
  }
  XMX(bck, L, BO_XC_SCALE) = sc;
  bck->totscale = log(sc);
  XMX(bck, L, BO_XC_E) = xE; XMX(bck, L, BO_XC_N) = xN; XMX(bck, L, BO_XC_J) = xJ;
  XMX(bck, L, BO_XC_B) = xB; XMX(bck, L, BO_XC_C) = xC;
  if (full) for (k = 0; k <= M; k++) {
    CELL(bck, L, k, BO_S_M) = k ? mc[k] : 0.0f; CELL(bck, L, k, BO_S_D) = k ? dc[k] : 0.0f; CELL(bck, L, k, BO_S_I) = k ? ic[k] : 0.0f; }

  for (i = L - 1; i >= 1; i--) {
    const float *rf = om->rfv + (size_t) dsq[i+1] * (M + 1);
    tmp = mp; mp = mc; mc = tmp;
    tmp = ip; ip = ic; ic = tmp;
    tmp = dpv; dpv = dc; dc = tmp;
    /* phase 1 (:556-592): mpv(k+1) = M(i+1,k+1) e(k+1) */
    xB = 0.0f;
    for (k = M; k >= 1; k--) {
      float mnext = (k < M) ? mp[k+1] * rf[k+1] : 0.0f;
      ic[k] = ip[k] * TF(BO_T_II, k) + mnext * TF(BO_T_IM, k);
      dc[k] = mnext * TF(BO_T_DM, k);
      mc[k] = ip[k] * TF(BO_T_MI, k) + mnext * TF(BO_T_MM, k);
    }
    for (k = 1; k <= M; k++) xB += mp[k] * rf[k] * TF(BO_T_BM, k-1);
    /* phase 2 (:594-606) */
    xC =  xC * om->xf[BO_X_C][BO_O_LOOP];
    xJ = (xB * om->xf[BO_X_J][BO_O_MOVE]) + (xJ * om->xf[BO_X_J][BO_O_LOOP]);
    xN = (xB * om->xf[BO_X_N][BO_O_MOVE]) + (xN * om->xf[BO_X_N][BO_O_LOOP]);
    xE = (xC * om->xf[BO_X_E][BO_O_MOVE]) + (xJ * om->xf[BO_X_E][BO_O_LOOP]);
    /* phases 3-5 (:609-648) */
    dc[M+1] = 0.0f;
    for (k = M; k >= 1; k--) {
      float dnext = (k < M) ? dc[k+1] : 0.0f;
      dc[k] = dc[k] + (dnext * TF(BO_T_DD, k) + xE);
      mc[k] = mc[k] + xE + dnext * TF(BO_T_MD, k);
    }
    /* scaling (:650-682) */
    if (xB > 1.0e16) bck->has_own_scales = 1;
    if (bck->has_own_scales) sc = (xB > 1.0e4) ? xB : 1.0f;
    else                     sc = XMX(fwd, i, BO_XC_SCALE);
    XMX(bck, i, BO_XC_SCALE) = sc;
    if (sc > 1.0f) {
      float sf = 1.0 / sc;
      xE /= sc; xN /= sc; xJ /= sc; xB /= sc; xC /= sc;
      for (k = 1; k <= M; k++) { mc[k] *= sf; dc[k] *= sf; ic[k] *= sf; }
      bck->totscale += log(sc);
    }
    XMX(bck, i, BO_XC_E) = xE; XMX(bck, i, BO_XC_N) = xN; XMX(bck, i, BO_XC_J) = xJ;
    XMX(bck, i, BO_XC_B) = xB; XMX(bck, i, BO_XC_C) = xC;
    if (full) for (k = 0; k <= M; k++) {
      CELL(bck, i, k, BO_S_M) = k ? mc[k] : 0.0f; CELL(bck, i, k, BO_S_D) = k ? dc[k] : 0.0f; CELL(bck, i, k, BO_S_I) = k ? ic[k] : 0.0f; }
  }
  /* termination at row 0 (:697-720) */
  if (L >= 1) {
    const float *rf = om->rfv + (size_t) dsq[1] * (M + 1);
    xB = 0.0f;
    for (k = 1; k <= M; k++) xB += (mc[k] * rf[k]) * TF(BO_T_BM, k-1);
  } else xB = 0.0f;
  xN = (xB * om->xf[BO_X_N][BO_O_MOVE]) + (xN * om->xf[BO_X_N][BO_O_LOOP]);
  XMX(bck, 0, BO_XC_B) = xB; XMX(bck, 0, BO_XC_C) = 0.0f; XMX(bck, 0, BO_XC_J) = 0.0f;
  XMX(bck, 0, BO_XC_N) = xN; XMX(bck, 0, BO_XC_E) = 0.0f; XMX(bck, 0, BO_XC_SCALE) = 1.0f;
  if (full) for (k = 0; k <= M; k++) CELL(bck, 0, k, BO_S_M) = CELL(bck, 0, k, BO_S_D) = CELL(bck, 0, k, BO_S_I) = 0.0f;
  free(mem);
  if (isnan(xN))          return BO_ERANGE;
  if (L > 0 && xN == 0.0) { if (opt_sc) *opt_sc = -INFINITY; return BO_ERANGE; }
  if (isinf(xN))          return BO_ERANGE;
  if (opt_sc) *opt_sc = bck->totscale + log(xN);
  return BO_OK;
}

/* p7_Decoding (decoding.c:76-139); pp may alias oxb */
int bo_Decoding(const BO_OPROFILE *om, const BO_MX *oxf, BO_MX *oxb, BO_MX *pp)
{
  int   L = oxf->L, M = om->M, i, k;
  float scaleproduct = 1.0 / XMX(oxb, 0, BO_XC_N);
  pp->M = M; pp->L = L;
  for (i = 1; i <= L; i++) {
    float totr = scaleproduct * XMX(oxf, i, BO_XC_SCALE);
    float bN = XMX(oxb, i, BO_XC_N), bJ = XMX(oxb, i, BO_XC_J), bC = XMX(oxb, i, BO_XC_C), bS = XMX(oxb, i, BO_XC_SCALE);
    for (k = 1; k <= M; k++) {
      CELL(pp, i, k, BO_S_M) = (CELL(oxf, i, k, BO_S_M) * CELL(oxb, i, k, BO_S_M)) * totr;
      CELL(pp, i, k, BO_S_D) = 0.0f;
      CELL(pp, i, k, BO_S_I) = (CELL(oxf, i, k, BO_S_I) * CELL(oxb, i, k, BO_S_I)) * totr;
    }
    CELL(pp, i, 0, BO_S_M) = CELL(pp, i, 0, BO_S_D) = CELL(pp, i, 0, BO_S_I) = 0.0f;
    XMX(pp, i, BO_XC_E) = 0.0f;
    XMX(pp, i, BO_XC_N) = XMX(oxf, i-1, BO_XC_N) * bN * om->xf[BO_X_N][BO_O_LOOP] * scaleproduct;
    XMX(pp, i, BO_XC_J) = XMX(oxf, i-1, BO_XC_J) * bJ * om->xf[BO_X_J][BO_O_LOOP] * scaleproduct;
    XMX(pp, i, BO_XC_C) = XMX(oxf, i-1, BO_XC_C) * bC * om->xf[BO_X_C][BO_O_LOOP] * scaleproduct;
    XMX(pp, i, BO_XC_B) = 0.0f;
    if (oxb->has_own_scales) scaleproduct *= XMX(oxf, i, BO_XC_SCALE) / bS;
  }
  for (k = 0; k <= M; k++) CELL(pp, 0, k, BO_S_M) = CELL(pp, 0, k, BO_S_D) = CELL(pp, 0, k, BO_S_I) = 0.0f;
  XMX(pp, 0, BO_XC_E) = XMX(pp, 0, BO_XC_N) = XMX(pp, 0, BO_XC_J) = XMX(pp, 0, BO_XC_C) = XMX(pp, 0, BO_XC_B) = 0.0f;
  return isinf(scaleproduct) ? BO_ERANGE : BO_OK;
}

/* p7_DomainDecoding (decoding.c:160-196); the three loop odds are om->xf[N|J|C][LOOP] at call time */
int bo_DomainDecoding(const float xf_loop_NJC[3], const BO_MX *oxf, const BO_MX *oxb, int own_scales,
                      float *btot, float *etot, float *mocc)
{
  int   L = oxf->L, i;
  float scaleproduct = 1.0 / XMX(oxb, 0, BO_XC_N);
  float njcp;
  btot[0] = etot[0] = mocc[0] = 0.0f;
  for (i = 1; i <= L; i++) {
    btot[i] = btot[i-1] + (XMX(oxf, i-1, BO_XC_B) * XMX(oxb, i-1, BO_XC_B) * XMX(oxf, i-1, BO_XC_SCALE) * scaleproduct);
    if (own_scales) scaleproduct *= XMX(oxf, i-1, BO_XC_SCALE) / XMX(oxb, i-1, BO_XC_SCALE);
    etot[i] = etot[i-1] + (XMX(oxf, i, BO_XC_E) * XMX(oxb, i, BO_XC_E) * XMX(oxf, i, BO_XC_SCALE) * scaleproduct);
    njcp  = XMX(oxf, i-1, BO_XC_N) * XMX(oxb, i, BO_XC_N) * xf_loop_NJC[0] * scaleproduct;
    njcp += XMX(oxf, i-1, BO_XC_J) * XMX(oxb, i, BO_XC_J) * xf_loop_NJC[1] * scaleproduct;
    njcp += XMX(oxf, i-1, BO_XC_C) * XMX(oxb, i, BO_XC_C) * xf_loop_NJC[2] * scaleproduct;
    mocc[i] = 1. - njcp;
  }
  return isinf(scaleproduct) ? BO_ERANGE : BO_OK;
}

/* p7_OptimalAccuracy (optacc.c:58-174): transitions act as masks (t > 0 ? value : 0.0 -- the SIMD `and` leaves +0.0,
 * not -inf, for a forbidden path) */
#define MASK(t, v) (((t) > 0.0f) ? (v) : 0.0f)
int bo_OptimalAccuracy(const BO_OPROFILE *om, const BO_MX *pp, BO_MX *ox, float *ret_e)
{
  int   M = om->M, L = pp->L, i, k;
  float t1, t2;
  ox->M = M; ox->L = L;
  for (k = 0; k <= M; k++) CELL(ox, 0, k, BO_S_M) = CELL(ox, 0, k, BO_S_I) = CELL(ox, 0, k, BO_S_D) = -INFINITY;
  XMX(ox, 0, BO_XC_E) = -INFINITY; XMX(ox, 0, BO_XC_N) = 0.; XMX(ox, 0, BO_XC_J) = -INFINITY;
  XMX(ox, 0, BO_XC_B) = 0.;        XMX(ox, 0, BO_XC_C) = -INFINITY;
  for (i = 1; i <= L; i++) {
    float xE = -INFINITY, xB = XMX(ox, i-1, BO_XC_B);
    CELL(ox, i, 0, BO_S_M) = CELL(ox, i, 0, BO_S_I) = CELL(ox, i, 0, BO_S_D) = -INFINITY;
    for (k = 1; k <= M; k++) {
      float sv = MASK(TF(BO_T_BM, k-1), xB);
      sv = fmaxf(sv, MASK(TF(BO_T_MM, k-1), CELL(ox, i-1, k-1, BO_S_M)));
      sv = fmaxf(sv, MASK(TF(BO_T_IM, k-1), CELL(ox, i-1, k-1, BO_S_I)));
      sv = fmaxf(sv, MASK(TF(BO_T_DM, k-1), CELL(ox, i-1, k-1, BO_S_D)));
      sv = sv + CELL(pp, i, k, BO_S_M);
      xE = fmaxf(xE, sv);
      CELL(ox, i, k, BO_S_M) = sv;
      sv = MASK(TF(BO_T_MI, k), CELL(ox, i-1, k, BO_S_M));
      sv = fmaxf(sv, MASK(TF(BO_T_II, k), CELL(ox, i-1, k, BO_S_I)));
      CELL(ox, i, k, BO_S_I) = sv + CELL(pp, i, k, BO_S_I);
    }
    CELL(ox, i, 1, BO_S_D) = -INFINITY;
    for (k = 2; k <= M; k++)
      CELL(ox, i, k, BO_S_D) = fmaxf(MASK(TF(BO_T_MD, k-1), CELL(ox, i, k-1, BO_S_M)), MASK(TF(BO_T_DD, k-1), CELL(ox, i, k-1, BO_S_D)));
    for (k = 1; k <= M; k++) xE = fmaxf(xE, CELL(ox, i, k, BO_S_D));
    XMX(ox, i, BO_XC_E) = xE;
    t1 = (om->xf[BO_X_J][BO_O_LOOP] == 0.0f) ? 0.0f : XMX(ox, i-1, BO_XC_J) + XMX(pp, i, BO_XC_J);
    t2 = (om->xf[BO_X_E][BO_O_LOOP] == 0.0f) ? 0.0f : XMX(ox, i, BO_XC_E);
    XMX(ox, i, BO_XC_J) = (t1 > t2) ? t1 : t2;
    t1 = (om->xf[BO_X_C][BO_O_LOOP] == 0.0f) ? 0.0f : XMX(ox, i-1, BO_XC_C) + XMX(pp, i, BO_XC_C);
    t2 = (om->xf[BO_X_E][BO_O_MOVE] == 0.0f) ? 0.0f : XMX(ox, i, BO_XC_E);
    XMX(ox, i, BO_XC_C) = (t1 > t2) ? t1 : t2;
    XMX(ox, i, BO_XC_N) = (om->xf[BO_X_N][BO_O_LOOP] == 0.0f) ? 0.0f : XMX(ox, i-1, BO_XC_N) + XMX(pp, i, BO_XC_N);
    t1 = (om->xf[BO_X_N][BO_O_MOVE] == 0.0f) ? 0.0f : XMX(ox, i, BO_XC_N);
    t2 = (om->xf[BO_X_J][BO_O_MOVE] == 0.0f) ? 0.0f : XMX(ox, i, BO_XC_J);
    XMX(ox, i, BO_XC_B) = (t1 > t2) ? t1 : t2;
  }
  *ret_e = XMX(ox, L, BO_XC_C);
  return BO_OK;
}

/* The D row above folds the reference's four SIMD passes into one serial max chain; at the first node of a stripe segment the
 * reference shifts in -inf (optacc.c:129), which the mask turns into 0.0 exactly as for any other forbidden path: the chain
 * D(k) = max(mask(tMD(k-1)) M(k-1), mask(tDD(k-1)) D(k-1)) is what every pass extends. */

/* p7_OATrace and its select_* helpers (optacc.c:225-425).  lanes = floats per SIMD vector of the CPU build (4 for SSE):
 * select_e scans cells in striped order, which decides ties. */
static int select_m(const BO_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  int   M = om->M, best = 0, z;
  float path[4];
  int   state[4] = { BO_ST_M, BO_ST_I, BO_ST_D, BO_ST_B };
  /* at k == 1 the shifted-in previous cells are 0.0 (rightshiftz), tested against transitions out of node 0 */
  float mpv = (k > 1) ? CELL(ox, i-1, k-1, BO_S_M) : 0.0f;
  float ipv = (k > 1) ? CELL(ox, i-1, k-1, BO_S_I) : 0.0f;
  float dpv = (k > 1) ? CELL(ox, i-1, k-1, BO_S_D) : 0.0f;
  path[3] = (TF(BO_T_BM, k-1) == 0.0f) ? -INFINITY : XMX(ox, i-1, BO_XC_B);
  path[0] = (TF(BO_T_MM, k-1) == 0.0f) ? -INFINITY : mpv;
  path[1] = (TF(BO_T_IM, k-1) == 0.0f) ? -INFINITY : ipv;
  path[2] = (TF(BO_T_DM, k-1) == 0.0f) ? -INFINITY : dpv;
  for (z = 1; z < 4; z++) if (path[z] > path[best]) best = z;
  return state[best];
}
static int select_d(const BO_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  int   M = om->M;
  float mpv = (k > 1) ? CELL(ox, i, k-1, BO_S_M) : 0.0f;
  float dpv = (k > 1) ? CELL(ox, i, k-1, BO_S_D) : 0.0f;
  float p0 = (TF(BO_T_MD, k-1) == 0.0f) ? -INFINITY : mpv;
  float p1 = (TF(BO_T_DD, k-1) == 0.0f) ? -INFINITY : dpv;
  return (p0 >= p1) ? BO_ST_M : BO_ST_D;
}
static int select_i(const BO_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  int   M = om->M;
  float p0 = (TF(BO_T_MI, k) == 0.0f) ? -INFINITY : CELL(ox, i-1, k, BO_S_M);
  float p1 = (TF(BO_T_II, k) == 0.0f) ? -INFINITY : CELL(ox, i-1, k, BO_S_I);
  return (p0 >= p1) ? BO_ST_M : BO_ST_I;
}
static int select_c(const BO_OPROFILE *om, const BO_MX *pp, const BO_MX *ox, int i)
{
  float p0 = (om->xf[BO_X_C][BO_O_LOOP] == 0.0f) ? -INFINITY : XMX(ox, i-1, BO_XC_C) + XMX(pp, i, BO_XC_C);
  float p1 = (om->xf[BO_X_E][BO_O_MOVE] == 0.0f) ? -INFINITY : XMX(ox, i, BO_XC_E);
  return (p0 > p1) ? BO_ST_C : BO_ST_E;
}
static int select_j(const BO_OPROFILE *om, const BO_MX *pp, const BO_MX *ox, int i)
{
  float p0 = (om->xf[BO_X_J][BO_O_LOOP] == 0.0f) ? -INFINITY : XMX(ox, i-1, BO_XC_J) + XMX(pp, i, BO_XC_J);
  float p1 = (om->xf[BO_X_E][BO_O_LOOP] == 0.0f) ? -INFINITY : XMX(ox, i, BO_XC_E);
  return (p0 > p1) ? BO_ST_J : BO_ST_E;
}
static int select_e(const BO_OPROFILE *om, const BO_MX *ox, int i, int lanes, int *ret_k)
{
  int   M = om->M, Q = (M - 1) / lanes + 1, q, r, k;
  float max = -INFINITY;
  int   smax = -1, kmax = 0;
  if (Q < 2) Q = 2;
  for (q = 0; q < Q; q++) {
    for (r = 0; r < lanes; r++) { k = r * Q + q + 1; float v = (k <= M) ? CELL(ox, i, k, BO_S_M) : -INFINITY;   /* pad cells hold -inf */
      if (v >= max) { max = v; smax = BO_ST_M; kmax = k; } }
    for (r = 0; r < lanes; r++) { k = r * Q + q + 1; float v = (k <= M) ? CELL(ox, i, k, BO_S_D) : -INFINITY;
      if (v > max)  { max = v; smax = BO_ST_D; kmax = k; } }
  }
  *ret_k = kmax;
  return smax;
}
static int select_b(const BO_OPROFILE *om, const BO_MX *ox, int i)
{
  float p0 = (om->xf[BO_X_N][BO_O_MOVE] == 0.0f) ? -INFINITY : XMX(ox, i, BO_XC_N);
  float p1 = (om->xf[BO_X_J][BO_O_MOVE] == 0.0f) ? -INFINITY : XMX(ox, i, BO_XC_J);
  return (p0 > p1) ? BO_ST_N : BO_ST_J;
}
static float get_postprob(const BO_MX *pp, int scur, int sprv, int k, int i)
{
  switch (scur) {
  case BO_ST_M: return CELL(pp, i, k, BO_S_M);
  case BO_ST_I: return CELL(pp, i, k, BO_S_I);
  case BO_ST_N: if (sprv == scur) return XMX(pp, i, BO_XC_N);   /* the reference falls through case labels here too */
  case BO_ST_C: if (sprv == scur) return XMX(pp, i, BO_XC_C);
  case BO_ST_J: if (sprv == scur) return XMX(pp, i, BO_XC_J);
  default:      return 0.0f;
  }
}

/* p7_trace_AppendWithPP + p7_trace_Reverse semantics are the frameshift container's with c = 0 (mx.c) */
int bo_OATrace(const BO_OPROFILE *om, const BO_MX *pp, const BO_MX *ox, int lanes, BO_TRACE *tr)
{
  int   i = ox->L, k = 0, s0, s1, status;
  float postprob;
  if (tr->N != 0) return BO_EINVAL;
  if ((status = bo_trace_append(tr, BO_ST_T, k, i, 0, 0.0f)) != BO_OK) return status;
  if ((status = bo_trace_append(tr, BO_ST_C, k, i, 0, 0.0f)) != BO_OK) return status;
  s0 = BO_ST_C;
  while (s0 != BO_ST_S) {
    switch (s0) {
    case BO_ST_M: s1 = select_m(om, ox, i, k); k--; i--; break;
    case BO_ST_D: s1 = select_d(om, ox, i, k); k--;      break;
    case BO_ST_I: s1 = select_i(om, ox, i, k);      i--; break;
    case BO_ST_N: s1 = (i == 0) ? BO_ST_S : BO_ST_N;     break;
    case BO_ST_C: s1 = select_c(om, pp, ox, i);          break;
    case BO_ST_J: s1 = select_j(om, pp, ox, i);          break;
    case BO_ST_E: s1 = select_e(om, ox, i, lanes, &k);   break;
    case BO_ST_B: s1 = select_b(om, ox, i);              break;
    default: return BO_EINVAL;
    }
    if (s1 == -1 || i < 0 || k < 0) return BO_EINVAL;
    postprob = get_postprob(pp, s1, s0, k, i);
    if ((status = bo_trace_append(tr, (char) s1, k, i, 0, postprob)) != BO_OK) return status;
    if ((s1 == BO_ST_N || s1 == BO_ST_J || s1 == BO_ST_C) && s1 == s0) i--;
    s0 = s1;
  }
  tr->M = om->M;
  tr->L = ox->L;
  bo_trace_reverse(tr);
  return BO_OK;
}

/* p7_Null2_ByExpectation (null2.c:44-125) */
int bo_Null2_ByExpectation(const BO_OPROFILE *om, const BO_MX *pp, float *null2)
{
  int    M = om->M, Ld = pp->L, i, k, x;
  float *em = calloc((size_t) 2 * (M + 1), sizeof(float)), *ei;
  float  xN, xC, xJ, norm, xfactor;
  if (!em) return BO_EMEM;
  ei = em + (M + 1);
  for (k = 1; k <= M; k++) { em[k] = CELL(pp, 1, k, BO_S_M); ei[k] = CELL(pp, 1, k, BO_S_I); }
  xN = XMX(pp, 1, BO_XC_N); xC = XMX(pp, 1, BO_XC_C); xJ = XMX(pp, 1, BO_XC_J);
  for (i = 2; i <= Ld; i++) {
    for (k = 1; k <= M; k++) { em[k] = CELL(pp, i, k, BO_S_M) + em[k]; ei[k] = CELL(pp, i, k, BO_S_I) + ei[k]; }
    xN += XMX(pp, i, BO_XC_N); xC += XMX(pp, i, BO_XC_C); xJ += XMX(pp, i, BO_XC_J);
  }
  norm = 1.0 / (float) Ld;
  for (k = 1; k <= M; k++) { em[k] *= norm; ei[k] *= norm; }
  xN *= norm; xC *= norm; xJ *= norm;
  xfactor = xN + xC + xJ;
  for (x = 0; x < BO_K; x++) {
    const float *rf = om->rfv + (size_t) x * (M + 1);
    float sv = 0.0f;
    for (k = 1; k <= M; k++) { sv += em[k] * rf[k]; sv += ei[k]; }
    null2[x] = sv + xfactor;
  }
  bo_abc_FAvgScVec(null2);
  null2[BO_K] = 1.0f; null2[BO_KP-2] = 1.0f; null2[BO_KP-1] = 1.0f;
  free(em);
  return BO_OK;
}
