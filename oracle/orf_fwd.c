/* orf_fwd.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 * Protein Forward parser over an ORF: p7_ForwardParser (src/impl_sse/fwdback.c:132; forward_engine :256-466),
 * probability space with sparse rescaling, un-striped, k-ordered; the D->D chain is the plain serial
 * recurrence the SIMD passes converge to (:352-395). */
#include <stdlib.h>
#include <math.h>
#include "bath_oracle.h"

#define TF(t,k) (om->tfv[(size_t)(t) * (M+1) + (k)])

int bo_ForwardParser(const uint8_t *dsq, int L, const BO_OPROFILE *om, float *opt_sc)
{
  int    M = om->M, i, k;
  float *mp, *ip, *dp, *mc, *ic, *dc, *tmp;
  float  xN, xE, xB, xC, xJ;
  double totscale = 0.0;
  float *mem = calloc((size_t) 6 * (M + 2), sizeof(float));
  if (!mem) return BO_EMEM;
  mp = mem; ip = mem + (M + 2); dp = mem + 2 * (M + 2); mc = mem + 3 * (M + 2); ic = mem + 4 * (M + 2); dc = mem + 5 * (M + 2);

  xE = 0.; xN = 1.; xJ = 0.; xC = 0.;
  xB = om->xf[BO_X_N][BO_O_MOVE];
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
      totscale += log(xE);
      xE = 1.0;
    }
    tmp = mp; mp = mc; mc = tmp;
    tmp = ip; ip = ic; ic = tmp;
    tmp = dp; dp = dc; dc = tmp;
  }
  free(mem);
  if (isnan(xC))             return BO_ERANGE;
  if (L > 0 && xC == 0.0)    { if (opt_sc) *opt_sc = -INFINITY; return BO_ERANGE; }
  if (isinf(xC))             return BO_ERANGE;
  if (opt_sc) *opt_sc = (float) totscale + log(xC * om->xf[BO_X_C][BO_O_MOVE]);
  return BO_OK;
}
