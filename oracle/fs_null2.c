/* fs_null2.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 * Follows src/impl_sse/null2_fs.c:53-139 (p7_Null2_fs_ByExpectation), un-striped.
 * Row 0 of <pp> is used as scratch exactly as the reference does. */
#include <math.h>
#include "bath_oracle.h"

#define RF(c,k) (om->rfv[(size_t)(c) * (M+1) + (k)])
#define XMX(mx,i,s) ((mx)->xmx[(size_t)(i) * BO_NXCELLS + (s)])
#define PCELL(mx,i,k,s) ((mx)->dp[((size_t)(i) * (M+1) + (k)) * BO_NSCELLS_FS + (s)])

int bo_Null2_fs_ByExpectation(const BO_FS_OPROFILE *om, BO_MX *pp, float *null2)
{
  int   M = om->M, Ld = pp->L;
  int   amino_offset = om->maxcodons;
  int   i, k, x, q, z, Q;
  float norm, xfactor;

  for (k = 1; k <= M; k++) {
    PCELL(pp, 0, k, BO_FS_M) = PCELL(pp, 1, k, BO_FS_M);
    PCELL(pp, 0, k, BO_FS_I) = PCELL(pp, 1, k, BO_FS_I);
  }
  XMX(pp, 0, BO_XC_N) = XMX(pp, 1, BO_XC_N);
  XMX(pp, 0, BO_XC_C) = XMX(pp, 1, BO_XC_C);
  XMX(pp, 0, BO_XC_J) = XMX(pp, 1, BO_XC_J);

  for (i = 2; i <= Ld; i++) {
    for (k = 1; k <= M; k++) {
      PCELL(pp, 0, k, BO_FS_M) = PCELL(pp, i, k, BO_FS_M) + PCELL(pp, 0, k, BO_FS_M);
      PCELL(pp, 0, k, BO_FS_I) = PCELL(pp, i, k, BO_FS_I) + PCELL(pp, 0, k, BO_FS_I);
    }
    XMX(pp, 0, BO_XC_N) += XMX(pp, i, BO_XC_N);
    XMX(pp, 0, BO_XC_C) += XMX(pp, i, BO_XC_C);
    XMX(pp, 0, BO_XC_J) += XMX(pp, i, BO_XC_J);
  }

  norm = 1.0 / (float) Ld;
  for (k = 1; k <= M; k++) {
    PCELL(pp, 0, k, BO_FS_M) = PCELL(pp, 0, k, BO_FS_M) * norm;
    PCELL(pp, 0, k, BO_FS_I) = PCELL(pp, 0, k, BO_FS_I) * norm;
  }
  XMX(pp, 0, BO_XC_N) *= norm;
  XMX(pp, 0, BO_XC_C) *= norm;
  XMX(pp, 0, BO_XC_J) *= norm;

  xfactor = XMX(pp, 0, BO_XC_N) + XMX(pp, 0, BO_XC_C) + XMX(pp, 0, BO_XC_J);
  Q = (M - 1) / 4 + 1; if (Q < 2) Q = 2;
  for (x = 0; x < BO_K; x++) {
    /* 4-lane striped accumulation + esl_sse_hsum_ps, as :118-127 */
    float lane[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
    for (q = 0; q < Q; q++)
      for (z = 0; z < 4; z++) {
        k = q + z * Q + 1;
        if (k <= M) {
          lane[z] = lane[z] + PCELL(pp, 0, k, BO_FS_M) * RF(amino_offset + x, k);
          lane[z] = lane[z] + PCELL(pp, 0, k, BO_FS_I);
        }
      }
    null2[x]  = (lane[0] + lane[2]) + (lane[1] + lane[3]);
    null2[x] += xfactor;
  }
  bo_abc_FAvgScVec(null2);
  null2[BO_K]      = 1.0;
  null2[BO_KP - 2] = 1.0;
  null2[BO_KP - 1] = 1.0;
  return BO_OK;
}
