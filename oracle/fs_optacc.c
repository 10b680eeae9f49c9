/* fs_optacc.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 * Optimal-accuracy fill and traceback for the frameshift model.
 * Follows src/impl_sse/optacc_fs.c:53-283 (fill) and :300-593 (traceback), un-striped.
 * Quirks kept on purpose:
 *  - transitions act as MASKS: a forbidden transition contributes 0.0, not -inf
 *    (_mm_and_ps with the t>0 mask, :140-181); a -inf source through an allowed
 *    transition stays -inf;
 *  - I(i,M) is forced to -inf (:207-212);
 *  - select_m compares predecessors at row i, column k-1 (:321-358), order M>I>D>B;
 *  - select_c/select_j early-outs i<4, i<=5 (:432,454); select_e first-k-wins (:477-484);
 *  - codon length = argmax of the five per-length posteriors (:504-518).
 */
#include <stdlib.h>
#include <math.h>
#include "bath_oracle.h"

#define TF(t,k) (om->tfv[(size_t)(t) * (M+1) + (k)])
#define XMX(mx,i,s) ((mx)->xmx[(size_t)(i) * BO_NXCELLS + (s)])
#define PCELL(mx,i,k,s) ((mx)->dp[((size_t)(i) * (M+1) + (k)) * BO_NSCELLS_FS + (s)])
#define OCELL(mx,i,k,s) ((mx)->dp[((size_t)(i) * (M+1) + (k)) * BO_NSCELLS    + (s)])

static inline float maskv(float t, float v) { return (t > 0.0f) ? v : 0.0f; }
static inline float fmax2(float a, float b) { return (a > b) ? a : b; }   /* _mm_max_ps(a,b): b if not a>b */

static int farg_max(const float *v, int n)
{
  int i, best = 0;
  for (i = 1; i < n; i++) if (v[i] > v[best]) best = i;
  return best;
}

int bo_OptimalAccuracy_Frameshift(const BO_FS_OPROFILE *om, const BO_MX *pp, BO_MX *ox, float *ret_e)
{
  int   M = om->M, L = pp->L;
  int   i, k, c;
  float xN, xE, xJ, xC, t1, t2;

  if (ox->nscells != BO_NSCELLS || ox->allocL < L || ox->M != M) return BO_EINVAL;
  ox->L = L;

  for (k = 0; k <= M; k++)
    OCELL(ox, 0, k, BO_S_M) = OCELL(ox, 0, k, BO_S_D) = OCELL(ox, 0, k, BO_S_I) = -INFINITY;
  XMX(ox, 0, BO_XC_E) = -INFINITY; XMX(ox, 0, BO_XC_N) = 0.0f; XMX(ox, 0, BO_XC_J) = -INFINITY;
  XMX(ox, 0, BO_XC_B) = 0.0f;      XMX(ox, 0, BO_XC_C) = -INFINITY;

  for (i = 1; i <= L; i++)
    {
      int   rows[6];
      float xBl[6];
      for (c = 1; c <= 5; c++) {
        rows[c] = (i >= c) ? i - c : 0;
        xBl[c]  = (i >= c) ? XMX(ox, i - c, BO_XC_B) : -INFINITY;
      }
      /* column 0 behaves as the -inf shifted-in lane */
      OCELL(ox, i, 0, BO_S_M) = OCELL(ox, i, 0, BO_S_D) = OCELL(ox, i, 0, BO_S_I) = -INFINITY;

      xE = -INFINITY;
      for (k = 1; k <= M; k++) {
        float bm = TF(BO_T_BM, k-1), mm = TF(BO_T_MM, k-1), im = TF(BO_T_IM, k-1), dm = TF(BO_T_DM, k-1);
        float svc[6], sv;
        for (c = 1; c <= 5; c++) {
          float s = maskv(bm, xBl[c]);
          s = fmax2(s, maskv(mm, OCELL(ox, rows[c], k-1, BO_S_M)));
          s = fmax2(s, maskv(im, OCELL(ox, rows[c], k-1, BO_S_I)));
          s = fmax2(s, maskv(dm, OCELL(ox, rows[c], k-1, BO_S_D)));
          svc[c] = s + PCELL(pp, i, k, BO_FS_M + c);
        }
        sv = fmax2(fmax2(svc[1], svc[2]), fmax2(fmax2(svc[3], svc[4]), svc[5]));
        xE = fmax2(xE, sv);
        OCELL(ox, i, k, BO_S_M) = sv;

        sv = maskv(TF(BO_T_MI, k), OCELL(ox, rows[3], k, BO_S_M));
        sv = fmax2(sv, maskv(TF(BO_T_II, k), OCELL(ox, rows[3], k, BO_S_I)));
        OCELL(ox, i, k, BO_S_I) = sv + PCELL(pp, i, k, BO_FS_I);
      }
      OCELL(ox, i, M, BO_S_I) = -INFINITY;

      OCELL(ox, i, 1, BO_S_D) = -INFINITY;
      for (k = 2; k <= M; k++) {
        float d = maskv(TF(BO_T_MD, k-1), OCELL(ox, i, k-1, BO_S_M));
        d = fmax2(maskv(TF(BO_T_DD, k-1), OCELL(ox, i, k-1, BO_S_D)), d);
        OCELL(ox, i, k, BO_S_D) = d;
      }
      for (k = 1; k <= M; k++) xE = fmax2(xE, OCELL(ox, i, k, BO_S_D));
      XMX(ox, i, BO_XC_E) = xE;

      if (i > 2) xN = (om->xf[BO_X_N][BO_O_LOOP] == 0.0f) ? 0.0f : XMX(ox, i-3, BO_XC_N) + XMX(pp, i, BO_XC_N);
      else       xN = (om->xf[BO_X_N][BO_O_LOOP] == 0.0f) ? 0.0f : XMX(pp, i, BO_XC_N);
      XMX(ox, i, BO_XC_N) = xN;

      if (i > 2) {
        t1 = (om->xf[BO_X_J][BO_O_LOOP] == 0.0f) ? 0.0f : XMX(ox, i-3, BO_XC_J) + XMX(pp, i, BO_XC_J);
        t2 = (om->xf[BO_X_E][BO_O_LOOP] == 0.0f) ? 0.0f : xE;
        xJ = (t1 > t2) ? t1 : t2;
      } else xJ = (om->xf[BO_X_E][BO_O_LOOP] == 0.0f) ? 0.0f : xE;
      XMX(ox, i, BO_XC_J) = xJ;

      if (i > 2) {
        t1 = (om->xf[BO_X_C][BO_O_LOOP] == 0.0f) ? 0.0f : XMX(ox, i-3, BO_XC_C) + XMX(pp, i, BO_XC_C);
        t2 = (om->xf[BO_X_E][BO_O_MOVE] == 0.0f) ? 0.0f : xE;
        xC = (t1 > t2) ? t1 : t2;
      } else xC = (om->xf[BO_X_E][BO_O_MOVE] == 0.0f) ? 0.0f : xE;
      XMX(ox, i, BO_XC_C) = xC;

      t1 = (om->xf[BO_X_N][BO_O_MOVE] == 0.0f) ? 0.0f : xN;
      t2 = (om->xf[BO_X_J][BO_O_MOVE] == 0.0f) ? 0.0f : xJ;
      XMX(ox, i, BO_XC_B) = (t1 > t2) ? t1 : t2;
    }

  *ret_e = XMX(ox, L, BO_XC_C) + XMX(ox, L-1, BO_XC_C) + XMX(ox, L-2, BO_XC_C);
  return BO_OK;
}

/* ---- traceback ---- */

static float get_postprob(const BO_MX *pp, int M, int scur, int sprv, int k, int i)
{
  switch (scur) {
  case BO_ST_M: return PCELL(pp, i, k, BO_FS_M);
  case BO_ST_I: return PCELL(pp, i, k, BO_FS_I);
  case BO_ST_N: if (sprv == scur) return XMX(pp, i, BO_XC_N); break;
  case BO_ST_C: if (sprv == scur) return XMX(pp, i, BO_XC_C); break;
  case BO_ST_J: if (sprv == scur) return XMX(pp, i, BO_XC_J); break;
  default: break;
  }
  return 0.0f;
}

static int select_m(const BO_FS_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  int   M = om->M;
  float path[4];
  static const int state[4] = { BO_ST_M, BO_ST_I, BO_ST_D, BO_ST_B };
  path[3] = (TF(BO_T_BM, k-1) == 0.0f) ? -INFINITY : XMX(ox, i, BO_XC_B);
  path[0] = (TF(BO_T_MM, k-1) == 0.0f) ? -INFINITY : OCELL(ox, i, k-1, BO_S_M);
  path[1] = (TF(BO_T_IM, k-1) == 0.0f) ? -INFINITY : OCELL(ox, i, k-1, BO_S_I);
  path[2] = (TF(BO_T_DM, k-1) == 0.0f) ? -INFINITY : OCELL(ox, i, k-1, BO_S_D);
  return state[farg_max(path, 4)];
}

static int select_d(const BO_FS_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  int   M = om->M;
  float path[2];
  path[0] = (TF(BO_T_MD, k-1) == 0.0f) ? -INFINITY : OCELL(ox, i, k-1, BO_S_M);
  path[1] = (TF(BO_T_DD, k-1) == 0.0f) ? -INFINITY : OCELL(ox, i, k-1, BO_S_D);
  return (path[0] >= path[1]) ? BO_ST_M : BO_ST_D;
}

static int select_i(const BO_FS_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  int   M = om->M;
  int   prev_i = (i >= 3) ? i - 3 : 0;
  float path[2];
  path[0] = (TF(BO_T_MI, k) == 0.0f) ? -INFINITY : OCELL(ox, prev_i, k, BO_S_M);
  path[1] = (TF(BO_T_II, k) == 0.0f) ? -INFINITY : OCELL(ox, prev_i, k, BO_S_I);
  return (path[0] >= path[1]) ? BO_ST_M : BO_ST_I;
}

static int select_c(const BO_FS_OPROFILE *om, const BO_MX *pp, const BO_MX *ox, int i)
{
  int   L = ox->L;
  float t1 = (om->xf[BO_X_C][BO_O_LOOP] == 0.0f) ? 0.0f : 1.0f;
  float t2 = (om->xf[BO_X_E][BO_O_MOVE] == 0.0f) ? 0.0f : 1.0f;
  float path[4];
  static const int state[4] = { BO_ST_C, BO_ST_C, BO_ST_C, BO_ST_E };
  if (i < 4) return BO_ST_E;
  path[0] = (t1 == 0.0f) ? -INFINITY : XMX(ox, i-3, BO_XC_C) + XMX(pp, i, BO_XC_C);
  path[1] = (i < L     && t1 != 0.0f) ? XMX(ox, i-2, BO_XC_C) + XMX(pp, i+1, BO_XC_C) : -INFINITY;
  path[2] = (i < L - 1 && t1 != 0.0f) ? XMX(ox, i-1, BO_XC_C) + XMX(pp, i+2, BO_XC_C) : -INFINITY;
  path[3] = (t2 == 0.0f) ? -INFINITY : XMX(ox, i, BO_XC_E);
  return state[farg_max(path, 4)];
}

static int select_j(const BO_FS_OPROFILE *om, const BO_MX *pp, const BO_MX *ox, int i)
{
  float path[2];
  static const int state[2] = { BO_ST_J, BO_ST_E };
  if (i <= 5) return BO_ST_E;
  path[0] = (om->xf[BO_X_J][BO_O_LOOP] == 0.0f) ? -INFINITY : XMX(ox, i, BO_XC_J) + XMX(pp, i, BO_XC_J);
  path[1] = (om->xf[BO_X_E][BO_O_LOOP] == 0.0f) ? -INFINITY : XMX(ox, i, BO_XC_E);
  return state[farg_max(path, 2)];
}

static int select_e(const BO_FS_OPROFILE *om, const BO_MX *ox, int i, int *ret_k)
{
  int   M = om->M, k, smax = BO_ST_M, kmax = 1;
  float max = -INFINITY;
  for (k = 1; k <= M; k++) {
    if (OCELL(ox, i, k, BO_S_M) > max) { max = OCELL(ox, i, k, BO_S_M); smax = BO_ST_M; kmax = k; }
    if (OCELL(ox, i, k, BO_S_D) > max) { max = OCELL(ox, i, k, BO_S_D); smax = BO_ST_D; kmax = k; }
  }
  *ret_k = kmax;
  return smax;
}

static int select_b(const BO_FS_OPROFILE *om, const BO_MX *ox, int i)
{
  float path[2];
  path[0] = (om->xf[BO_X_N][BO_O_MOVE] == 0.0f) ? -INFINITY : XMX(ox, i, BO_XC_N);
  path[1] = (om->xf[BO_X_J][BO_O_MOVE] == 0.0f) ? -INFINITY : XMX(ox, i, BO_XC_J);
  return (path[0] > path[1]) ? BO_ST_N : BO_ST_J;
}

static int select_codon(const BO_MX *pp, int M, int i, int k)
{
  float codon[5];
  int c;
  for (c = 0; c < 5; c++) codon[c] = PCELL(pp, i, k, BO_FS_M + 1 + c);
  return farg_max(codon, 5) + 1;
}

/* optacc_fs.c:547-593 */
int bo_OATrace_Frameshift(const BO_FS_OPROFILE *om, const BO_MX *pp, const BO_MX *ox, BO_TRACE *tr)
{
  int   M = om->M;
  int   i = ox->L, k = 0, c = 0;
  int   sprv, scur, status;
  float postprob;

  if (tr->N != 0) return BO_EINVAL;
  if ((status = bo_trace_append(tr, BO_ST_T, k, i, c, 0.0f)) != BO_OK) return status;
  if ((status = bo_trace_append(tr, BO_ST_C, k, i, c, 0.0f)) != BO_OK) return status;

  sprv = BO_ST_C;
  while (sprv != BO_ST_S)
    {
      switch (sprv) {
      case BO_ST_M: scur = select_m(om, ox, i, k); k--;    break;
      case BO_ST_D: scur = select_d(om, ox, i, k); k--;    break;
      case BO_ST_I: scur = select_i(om, ox, i, k); i -= 3; break;
      case BO_ST_N: scur = (i == 0) ? BO_ST_S : BO_ST_N;   break;
      case BO_ST_C: scur = select_c(om, pp, ox, i);        break;
      case BO_ST_J: scur = select_j(om, pp, ox, i);        break;
      case BO_ST_E: scur = select_e(om, ox, i, &k);        break;
      case BO_ST_B: scur = select_b(om, ox, i);            break;
      default: return BO_EINVAL;
      }
      if (i < 0 || k < 0) return BO_EINVAL;   /* guard the oracle against walking off the matrix */

      postprob = get_postprob(pp, M, scur, sprv, k, i);
      c = (scur == BO_ST_M) ? select_codon(pp, M, i, k) : 0;
      if ((status = bo_trace_append(tr, (char) scur, k, i, c, postprob)) != BO_OK) return status;

      if ((scur == BO_ST_N || scur == BO_ST_C || scur == BO_ST_J) && scur == sprv) i--;
      sprv = scur;
      i   -= c;
    }
  tr->M = M;
  tr->L = ox->L;
  bo_trace_reverse(tr);
  return BO_OK;
}
