/* stotrace.c -- ORACLE (test infrastructure only; never linked into the product).
 *
 * The multi-domain branch of the standard-translation domain definition:
 *   p7_StochasticTrace                    src/impl_sse/stotrace.c:72-113, select_* :127-286
 *   p7_trace_Index                        src/p7_trace.c:2592-2625
 *   p7_Null2_ByTrace                      src/impl_sse/null2.c:131-219
 *   region_trace_ensemble                 src/p7_domaindef.c:766-860 (sampling, per-residue null2 scores, clustering, dominated clusters)
 * PARITY UNPINNED, for the reasons given in fs_stotrace.c (no shipped output exercises the branch; generator, FChoose, FNorm,
 * FAvgScVec, the horizontal sum and the clustering routine are Easel's, restated).  The matrix is un-striped, so select_e and
 * the null2 sums walk the nodes in the order the striped SSE loops visit them (q outer, lane r inner: k = r Q + q + 1). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "bath_oracle.h"

#define PC(mx,i,k,s) ((mx)->dp[((size_t)(i) * ((mx)->M + 1) + (k)) * BO_NSCELLS + (s)])
#define XM(mx,i,s)   ((mx)->xmx[(size_t)(i) * BO_NXCELLS + (s)])
#define TF(t,k)      (om->tfv[(size_t)(t) * (om->M + 1) + (k)])

void bo_fnorm(float *v, int n);
int  bo_fchoose(BO_RNG *r, const float *p, int n);

static int select_m(BO_RNG *r, const BO_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  static const int state[4] = { BO_ST_B, BO_ST_M, BO_ST_I, BO_ST_D };
  float path[4];
  path[0] = XM(ox, i - 1, BO_XC_B) * TF(BO_T_BM, k - 1);
  path[1] = (k > 1) ? PC(ox, i - 1, k - 1, BO_S_M) * TF(BO_T_MM, k - 1) : 0.0f;      /* node 0: the right shift brings in zeros */
  path[2] = (k > 1) ? PC(ox, i - 1, k - 1, BO_S_I) * TF(BO_T_IM, k - 1) : 0.0f;
  path[3] = (k > 1) ? PC(ox, i - 1, k - 1, BO_S_D) * TF(BO_T_DM, k - 1) : 0.0f;
  bo_fnorm(path, 4);
  return state[bo_fchoose(r, path, 4)];
}
static int select_d(BO_RNG *r, const BO_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  float path[2];
  path[0] = (k > 1) ? PC(ox, i, k - 1, BO_S_M) * TF(BO_T_MD, k - 1) : 0.0f;
  path[1] = (k > 1) ? PC(ox, i, k - 1, BO_S_D) * TF(BO_T_DD, k - 1) : 0.0f;
  bo_fnorm(path, 2);
  return bo_fchoose(r, path, 2) == 0 ? BO_ST_M : BO_ST_D;
}
static int select_i(BO_RNG *r, const BO_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  float path[2];
  path[0] = PC(ox, i - 1, k, BO_S_M) * TF(BO_T_MI, k);
  path[1] = PC(ox, i - 1, k, BO_S_I) * TF(BO_T_II, k);
  bo_fnorm(path, 2);
  return bo_fchoose(r, path, 2) == 0 ? BO_ST_M : BO_ST_I;
}
static int select_cj(BO_RNG *r, const BO_MX *ox, int i, int cell, float loop, float e_odds, int self)
{
  float path[2];
  if (i < 1) return BO_ST_E;
  path[0] = XM(ox, i - 1, cell) * loop;
  path[1] = XM(ox, i, BO_XC_E) * e_odds * XM(ox, i, BO_XC_SCALE);
  bo_fnorm(path, 2);
  return bo_fchoose(r, path, 2) == 0 ? self : BO_ST_E;
}
static int select_e(BO_RNG *r, const BO_MX *ox, int i, int *ret_k)
{
  const int M = ox->M, Q = (M - 1) / 4 + 1 > 2 ? (M - 1) / 4 + 1 : 2;
  double sum = 0.0, roll = bo_random(r), norm = 1.0 / XM(ox, i, BO_XC_E);
  float  nf = (float) norm;
  int    q, z, pass;
  for (pass = 0; pass < 1000; pass++)
    for (q = 0; q < Q; q++) {
      for (z = 0; z < 4; z++) { int k = z * Q + q + 1; sum += (k <= M) ? PC(ox, i, k, BO_S_M) * nf : 0.0f; if (roll < sum) { *ret_k = k; return BO_ST_M; } }
      for (z = 0; z < 4; z++) { int k = z * Q + q + 1; sum += (k <= M) ? PC(ox, i, k, BO_S_D) * nf : 0.0f; if (roll < sum) { *ret_k = k; return BO_ST_D; } }
    }
  return -1;
}
static int select_b(BO_RNG *r, const BO_OPROFILE *om, const BO_MX *ox, int i)
{
  float path[2];
  path[0] = XM(ox, i, BO_XC_N) * om->xf[BO_X_N][BO_O_MOVE];
  path[1] = XM(ox, i, BO_XC_J) * om->xf[BO_X_J][BO_O_MOVE];
  bo_fnorm(path, 2);
  return bo_fchoose(r, path, 2) == 0 ? BO_ST_N : BO_ST_J;
}

int bo_StochasticTrace(BO_RNG *rng, int L, const BO_OPROFILE *om, const BO_MX *ox, BO_TRACE *tr)
{
  int i = L, k = 0, s0, s1;
  bo_trace_append(tr, BO_ST_T, k, i, 0, 0.0f);
  bo_trace_append(tr, BO_ST_C, k, i, 0, 0.0f);
  s0 = BO_ST_C;
  while (s0 != BO_ST_S) {
    switch (s0) {
    case BO_ST_M: if (i < 1 || k < 1) return BO_EINVAL; s1 = select_m(rng, om, ox, i, k); k--; i--; break;
    case BO_ST_D: s1 = select_d(rng, om, ox, i, k); k--;      break;
    case BO_ST_I: if (i < 1) return BO_EINVAL; s1 = select_i(rng, om, ox, i, k); i--; break;
    case BO_ST_N: s1 = (i == 0) ? BO_ST_S : BO_ST_N;          break;
    case BO_ST_C: s1 = select_cj(rng, ox, i, BO_XC_C, om->xf[BO_X_C][BO_O_LOOP], om->xf[BO_X_E][BO_O_MOVE], BO_ST_C); break;
    case BO_ST_J: s1 = select_cj(rng, ox, i, BO_XC_J, om->xf[BO_X_J][BO_O_LOOP], om->xf[BO_X_E][BO_O_LOOP], BO_ST_J); break;
    case BO_ST_E: s1 = select_e(rng, ox, i, &k);              break;
    case BO_ST_B: s1 = select_b(rng, om, ox, i);              break;
    default: return BO_EINVAL;
    }
    if (s1 == -1) return BO_EINVAL;
    bo_trace_append(tr, (char) s1, k, i, 0, 0.0f);
    if ((s1 == BO_ST_N || s1 == BO_ST_J || s1 == BO_ST_C) && s1 == s0) i--;
    s0 = s1;
    if (i < 0) return BO_EINVAL;
  }
  tr->M = om->M; tr->L = L;
  bo_trace_reverse(tr);
  return BO_OK;
}

/* p7_Null2_ByTrace over trace elements zstart..zend; null2[Kp] */
static void null2_by_trace(const BO_OPROFILE *om, const BO_TRACE *tr, int zstart, int zend, float *null2)
{
  const int M = om->M, Q = (M - 1) / 4 + 1 > 2 ? (M - 1) / 4 + 1 : 2;
  float *cnt = calloc((size_t) M + 1, sizeof(float));
  float  xN = 0.0f, xC = 0.0f, xJ = 0.0f, norm, xfactor;
  int    Ld = 0, z, x, q, r, k;
  static const int members[6][2] = { { 2, 11 }, { 7, 9 }, { 3, 13 }, { 8, 8 }, { 1, 1 }, { -1, -1 } };
  for (z = zstart; z <= zend; z++) {
    if (tr->i[z] == 0) continue;
    Ld++;
    if (tr->k[z] > 0) cnt[tr->k[z]] += 1.0f;            /* M or I: both land on the node's match cell (:163) */
    else switch (tr->st[z]) { case BO_ST_N: xN += 1.0f; break; case BO_ST_C: xC += 1.0f; break; case BO_ST_J: xJ += 1.0f; break; default: break; }
  }
  norm = 1.0f / (float) Ld;
  for (k = 1; k <= M; k++) cnt[k] *= norm;
  xN *= norm; xC *= norm; xJ *= norm;
  xfactor = xN + xC + xJ;
  for (x = 0; x < BO_K; x++) {
    float sv[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
    for (q = 0; q < Q; q++)
      for (r = 0; r < 4; r++) { k = r * Q + q + 1; if (k <= M) sv[r] = sv[r] + cnt[k] * om->rfv[(size_t) x * (M + 1) + k]; }
    null2[x] = (sv[0] + sv[1]) + (sv[2] + sv[3]);
    null2[x] += xfactor;
  }
  for (x = BO_K + 1; x <= BO_KP - 3; x++) {             /* esl_abc_FAvgScVec */
    const int *mb = members[x - BO_K - 1];
    float sum = 0.0f; int n = 0, y;
    for (y = 0; y < BO_K; y++) if (mb[0] < 0 || y == mb[0] || y == mb[1]) { sum += null2[y]; n++; }
    null2[x] = sum / (float) n;
  }
  null2[BO_K] = 1.0f; null2[BO_KP - 2] = 1.0f; null2[BO_KP - 1] = 1.0f;
  free(cnt);
}

/* region_trace_ensemble on a filled multihit Forward matrix of region ireg..jreg of the ORF dsq[1..]: sampled segments (ORF
 * coordinates) to samples[], consensus envelopes to out[], n2sc[ireg..jreg] = per-residue null2 scores.  Returns the number of
 * envelopes, -1 if a traceback fails. */
int bo_region_trace_ensemble(const BO_OPROFILE *om, const uint8_t *dsq, const BO_MX *fwd, int ireg, int jreg, uint32_t seed, int nsamples,
                             BO_SEGMENT *samples, int max_samples, int *ret_nsamples, BO_SEGMENT *out, int max_out, float *n2sc)
{
  const int Lr = jreg - ireg + 1;
  BO_RNG rng;
  BO_TRACE *tr = bo_trace_create();
  BO_SEGMENT *sp = malloc(sizeof(BO_SEGMENT) * (size_t) nsamples * 64);
  float null2[BO_KP];
  int t, z, n = 0, nc, pos;
  for (pos = ireg; pos <= jreg; pos++) n2sc[pos] = 0.0f;
  bo_rng_init(&rng, seed);
  for (t = 0; t < nsamples; t++) {
    int tfrom = -1, sqfrom = 0, sqto = 0, hmmfrom = 0, hmmto = 0;
    if (bo_StochasticTrace(&rng, Lr, om, fwd, tr) != BO_OK) { bo_trace_destroy(tr); free(sp); return -1; }
    pos = 1;
    for (z = 0; z < tr->N; z++)                          /* p7_trace_Index, one domain at a time */
      switch (tr->st[z]) {
      case BO_ST_B: tfrom = z; sqfrom = 0; hmmfrom = 0; break;
      case BO_ST_M:
        if (sqfrom == 0) sqfrom = tr->i[z];
        if (hmmfrom == 0) hmmfrom = tr->k[z];
        sqto = tr->i[z]; hmmto = tr->k[z];
        break;
      case BO_ST_E:
        if (n < nsamples * 64) { sp[n].idx = t; sp[n].i = sqfrom + ireg - 1; sp[n].j = sqto + ireg - 1; sp[n].k = hmmfrom; sp[n].m = hmmto; sp[n].prob = 0.0f; n++; }
        null2_by_trace(om, tr, tfrom, z, null2);
        for (; pos <= sqfrom; pos++) n2sc[ireg + pos - 1] += 1.0f;
        for (; pos <= sqto;   pos++) n2sc[ireg + pos - 1] += null2[dsq[ireg + pos - 1]];
        break;
      default: break;
      }
    for (; pos <= Lr; pos++) n2sc[ireg + pos - 1] += 1.0f;
    bo_trace_reuse(tr);
  }
  for (pos = ireg; pos <= jreg; pos++) n2sc[pos] = logf(n2sc[pos] / (float) nsamples);
  if (samples) for (z = 0; z < n && z < max_samples; z++) samples[z] = sp[z];
  if (ret_nsamples) *ret_nsamples = n;
  nc = bo_spensemble_Cluster(sp, n, nsamples, 0.8f, 1, 4, 0.25f, 0.02f, out, max_out);
  bo_trace_destroy(tr); free(sp);
  return nc;
}
