/* fs_decoding.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 * Posterior decoding for the frameshift model.
 * Follows src/impl_sse/decoding_fs.c:55-200 (p7_Decoding_Frameshift) and
 * :245-359 (p7_DomainDecoding_Frameshift), un-striped. */
#include <stdlib.h>
#include <math.h>
#include "bath_oracle.h"

#define XMX(mx,i,s) ((mx)->xmx[(size_t)(i) * BO_NXCELLS + (s)])
#define FCELL(mx,i,k,s) ((mx)->dp[((size_t)(i) * (M+1) + (k)) * BO_NSCELLS_FS + (s)])
#define BCELL(mx,i,k,s) ((mx)->dp[((size_t)(i) * (M+1) + (k)) * BO_NSCELLS    + (s)])

static int cumulative_scales(const BO_MX *oxf, const BO_MX *oxb, int L, float **ret_sf, float **ret_sb, float *ret_log_inv_Z)
{
  float *log_sfwd = malloc(sizeof(float) * (L + 2));
  float *log_sbck = malloc(sizeof(float) * (L + 2));
  int i;
  if (!log_sfwd || !log_sbck) { free(log_sfwd); free(log_sbck); return BO_EMEM; }

  log_sfwd[0] = logf(XMX(oxf, 0, BO_XC_SCALE));
  for (i = 1; i <= L; i++) log_sfwd[i] = log_sfwd[i-1] + logf(XMX(oxf, i, BO_XC_SCALE));
  log_sbck[L+1] = 0.0f;
  for (i = L; i >= 0; i--) log_sbck[i] = log_sbck[i+1] + logf(XMX(oxb, i, BO_XC_SCALE));

  *ret_log_inv_Z = -bo_FLogsum(logf(XMX(oxb, 0, BO_XC_N)) + log_sbck[0],
                      bo_FLogsum(logf(XMX(oxb, 1, BO_XC_N)) + log_sbck[1],
                                 logf(XMX(oxb, 2, BO_XC_N)) + log_sbck[2]));
  *ret_sf = log_sfwd; *ret_sb = log_sbck;
  return BO_OK;
}

/* decoding_fs.c:55-200.  <fwd> (8-cell) is overwritten with posteriors. */
int bo_Decoding_Frameshift(const BO_FS_OPROFILE *om, BO_MX *fwd, const BO_MX *bck)
{
  int    L = fwd->L, M = om->M;
  int    i, k, c, status;
  float *log_sfwd = NULL, *log_sbck = NULL, log_inv_Z;
  float  nlag[4], jlag[4], clag[4];
  const float N_odds = om->xf[BO_X_N][BO_O_LOOP];
  const float J_odds = om->xf[BO_X_J][BO_O_LOOP];
  const float C_odds = om->xf[BO_X_C][BO_O_LOOP];

  if ((status = cumulative_scales(fwd, bck, L, &log_sfwd, &log_sbck, &log_inv_Z)) != BO_OK) return status;

  nlag[0] = XMX(fwd, 0, BO_XC_N); jlag[0] = XMX(fwd, 0, BO_XC_J); clag[0] = XMX(fwd, 0, BO_XC_C);
  nlag[1] = nlag[2] = nlag[3] = 0.0f;
  jlag[1] = jlag[2] = jlag[3] = 0.0f;
  clag[1] = clag[2] = clag[3] = 0.0f;

  for (k = 0; k <= M; k++) for (c = 0; c < BO_NSCELLS_FS; c++) FCELL(fwd, 0, k, c) = 0.0f;
  XMX(fwd, 0, BO_XC_E) = XMX(fwd, 0, BO_XC_N) = XMX(fwd, 0, BO_XC_J) = XMX(fwd, 0, BO_XC_B) = XMX(fwd, 0, BO_XC_C) = 0.0f;

  for (i = 1; i <= L; i++)
    {
      float fN3, fJ3, fC3, bck_N, bck_J, bck_C, factor_mdi, raw_denom, N_pp, J_pp, C_pp, inv_denom, scv;
      float lane[4]; int Q = (M - 1) / 4 + 1; if (Q < 2) Q = 2;

      nlag[i % 4] = XMX(fwd, i, BO_XC_N);
      jlag[i % 4] = XMX(fwd, i, BO_XC_J);
      clag[i % 4] = XMX(fwd, i, BO_XC_C);
      fN3 = nlag[(i + 1) % 4]; fJ3 = jlag[(i + 1) % 4]; fC3 = clag[(i + 1) % 4];
      bck_N = XMX(bck, i, BO_XC_N); bck_J = XMX(bck, i, BO_XC_J); bck_C = XMX(bck, i, BO_XC_C);

      factor_mdi = expf(log_sfwd[i] + log_sbck[i] + log_inv_Z);
      if (isinf(factor_mdi)) { free(log_sfwd); free(log_sbck); return BO_ERANGE; }

      lane[0] = lane[1] = lane[2] = lane[3] = 0.0f;
      for (k = 1; k <= M; k++) {
        float bM = BCELL(bck, i, k, BO_S_M);
        float bI = BCELL(bck, i, k, BO_S_I);
        FCELL(fwd, i, k, BO_FS_D) = 0.0f;
        FCELL(fwd, i, k, BO_FS_I) = FCELL(fwd, i, k, BO_FS_I) * bI;
        for (c = 0; c < 6; c++) FCELL(fwd, i, k, BO_FS_M + c) = FCELL(fwd, i, k, BO_FS_M + c) * bM;
      }
      /* per-row denominator: 4-lane striped accumulation then hsum, as :152-160 */
      {
        int q, z;
        for (q = 0; q < Q; q++)
          for (z = 0; z < 4; z++) {
            k = q + z * Q + 1;
            if (k <= M) lane[z] = lane[z] + (FCELL(fwd, i, k, BO_FS_M) + FCELL(fwd, i, k, BO_FS_I));
          }
        raw_denom = (lane[0] + lane[2]) + (lane[1] + lane[3]);
      }

      if (i > 2) {
        float factor_njc = expf(log_sfwd[i-3] + log_sbck[i] + log_inv_Z);
        N_pp = fN3 * bck_N * N_odds * factor_njc;
        J_pp = fJ3 * bck_J * J_odds * factor_njc;
        C_pp = fC3 * bck_C * C_odds * factor_njc;
      } else {
        float factor_nsmall = expf(log_sbck[i] + log_inv_Z);
        N_pp = bck_N * factor_nsmall;
        J_pp = 0.0f;
        C_pp = 0.0f;
      }
      inv_denom = 1.0f / (raw_denom * factor_mdi + N_pp + J_pp + C_pp);
      if (isinf(inv_denom)) { free(log_sfwd); free(log_sbck); return BO_ERANGE; }

      scv = factor_mdi * inv_denom;
      for (k = 1; k <= M; k++)
        for (c = 0; c < BO_NSCELLS_FS; c++) FCELL(fwd, i, k, c) = FCELL(fwd, i, k, c) * scv;

      XMX(fwd, i, BO_XC_E) = 0.0f;
      XMX(fwd, i, BO_XC_B) = 0.0f;
      XMX(fwd, i, BO_XC_N) = N_pp * inv_denom;
      XMX(fwd, i, BO_XC_J) = J_pp * inv_denom;
      XMX(fwd, i, BO_XC_C) = C_pp * inv_denom;
    }

  free(log_sfwd); free(log_sbck);
  return BO_OK;
}

/* decoding_fs.c:245-359.  xf_loop_NJC = {tNL, tJL, tCL} of the profile the
 * reference passes (om_fs5 in p7_domaindef.c:320 -- NOT the om_fs3 the parsers ran with). */
int bo_DomainDecoding_Frameshift(const float xf_loop_NJC[3], const BO_MX *oxf, const BO_MX *oxb,
                                 float *btot, float *etot, float *mocc)
{
  int    L = oxf->L, i, status;
  float *log_sfwd = NULL, *log_sbck = NULL, log_inv_Z, njcp;
  const float tNL = xf_loop_NJC[0], tJL = xf_loop_NJC[1], tCL = xf_loop_NJC[2];

  if ((status = cumulative_scales(oxf, oxb, L, &log_sfwd, &log_sbck, &log_inv_Z)) != BO_OK) return status;

  btot[0] = btot[1] = btot[2] = 0.;
  etot[0] = etot[1] = etot[2] = 0.;
  mocc[0] = mocc[1] = mocc[2] = 0.;

  for (i = 3; i <= L; i++)
    {
      btot[i] = btot[i-3] + XMX(oxf, i-3, BO_XC_B) * XMX(oxb, i-3, BO_XC_B) * expf(log_sfwd[i-3] + log_sbck[i-3] + log_inv_Z);
      etot[i] = etot[i-3] + XMX(oxf, i,   BO_XC_E) * XMX(oxb, i,   BO_XC_E) * expf(log_sfwd[i]   + log_sbck[i]   + log_inv_Z);

      njcp = 0.;
      njcp += XMX(oxf, i-3, BO_XC_N) * XMX(oxb, i, BO_XC_N) * tNL * expf(log_sfwd[i-3] + log_sbck[i] + log_inv_Z);
      if (i < L)
        njcp += XMX(oxf, i-2, BO_XC_N) * XMX(oxb, i+1, BO_XC_N) * tNL * expf(log_sfwd[i-2] + log_sbck[i+1] + log_inv_Z);
      if (i < L - 1)
        njcp += XMX(oxf, i-1, BO_XC_N) * XMX(oxb, i+2, BO_XC_N) * tNL * expf(log_sfwd[i-1] + log_sbck[i+2] + log_inv_Z);

      njcp += XMX(oxf, i-3, BO_XC_J) * XMX(oxb, i, BO_XC_J) * tJL * expf(log_sfwd[i-3] + log_sbck[i] + log_inv_Z);
      if (i < L)
        njcp += XMX(oxf, i-2, BO_XC_J) * XMX(oxb, i+1, BO_XC_J) * tJL * expf(log_sfwd[i-2] + log_sbck[i+1] + log_inv_Z);
      if (i < L - 1)
        njcp += XMX(oxf, i-1, BO_XC_J) * XMX(oxb, i+2, BO_XC_J) * tJL * expf(log_sfwd[i-1] + log_sbck[i+2] + log_inv_Z);

      njcp += XMX(oxf, i-3, BO_XC_C) * XMX(oxb, i, BO_XC_C) * tCL * expf(log_sfwd[i-3] + log_sbck[i] + log_inv_Z);
      if (i < L)
        njcp += XMX(oxf, i-2, BO_XC_C) * XMX(oxb, i+1, BO_XC_C) * tCL * expf(log_sfwd[i-2] + log_sbck[i+1] + log_inv_Z);
      if (i < L - 1)
        njcp += XMX(oxf, i-1, BO_XC_C) * XMX(oxb, i+2, BO_XC_C) * tCL * expf(log_sfwd[i-1] + log_sbck[i+2] + log_inv_Z);

      mocc[i] = 1. - njcp;
    }
  free(log_sfwd); free(log_sbck);
  return BO_OK;
}
