/* fs_stotrace.c -- ORACLE (test infrastructure only; never linked into the product).
 *
 * Stochastic traceback over a frameshift Forward matrix and the clustering of the sampled domain
 * coordinates that splits a multi-domain region into envelopes:
 *   p7_StochasticTrace_Frameshift        src/impl_sse/stotrace_fs.c:72-128, select_* :150-365
 *   p7_trace_fs_Index                    src/p7_trace.c:2645-2680
 *   p7_spensemble_Add / _fs_Cluster      src/p7_spensemble.c:146-170, :226-256 (link rule), :498-640
 *   region_trace_ensemble_frameshift     src/p7_domaindef.c:892-954
 *
 * PARITY UNPINNED for this file: the reference ships no expected output that exercises the multi-domain branch, and
 * three pieces come from Easel, which is absent from /root/reference (TravisWheelerLab/easel, branch BATH, unpinned):
 *   - the "fast" generator behind esl_randomness_CreateFast (src/p7_pipeline.c:140): a 32-bit linear congruential
 *     generator x <- 69069 x + 1 seeded through Bob Jenkins' three-word mix, esl_random() = x / 2^32;
 *   - esl_rnd_FChoose (running double sum against one draw; on falling through, uniform draws until a non-zero entry) and
 *     esl_vec_FNorm (divide by the compensated sum; all-zero vectors become uniform);
 *   - esl_cluster_SingleLinkage (stack-based connected components over the link predicate).
 * They are restated from Easel's published sources as the author knows them; the matrix layout is un-striped, so
 * select_e walks the nodes in the order the striped SSE loop visits them (q outer, lane r inner: k = r*Q + q + 1). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "bath_oracle.h"

#define FC(mx,i,k,s) ((mx)->dp[((size_t)(i) * ((mx)->M + 1) + (k)) * BO_NSCELLS_FS + (s)])
#define XM(mx,i,s)   ((mx)->xmx[(size_t)(i) * BO_NXCELLS + (s)])
#define TF(t,k)      (om->tfv[(size_t)(t) * (om->M + 1) + (k)])

/* ---- Easel's fast generator ---- */
static uint32_t jenkins_mix3(uint32_t a, uint32_t b, uint32_t c)
{
  a -= b; a -= c; a ^= (c >> 13);
  b -= c; b -= a; b ^= (a << 8);
  c -= a; c -= b; c ^= (b >> 13);
  a -= b; a -= c; a ^= (c >> 12);
  b -= c; b -= a; b ^= (a << 16);
  c -= a; c -= b; c ^= (b >> 5);
  a -= b; a -= c; a ^= (c >> 3);
  b -= c; b -= a; b ^= (a << 10);
  c -= a; c -= b; c ^= (b >> 15);
  return c;
}

void bo_rng_init(BO_RNG *r, uint32_t seed)
{
  r->seed = seed;
  r->x = jenkins_mix3(seed, 87654321u, 12345678u);
  if (r->x == 0) r->x = 42;
}

double bo_random(BO_RNG *r)
{
  r->x = r->x * 69069u + 1u;
  return (double) r->x / 4294967296.0;
}

void bo_fnorm(float *v, int n)
{
  float sum = 0.0f, c = 0.0f, y, t;
  int   x;
  for (x = 0; x < n; x++) { y = v[x] - c; t = sum + y; c = (t - sum) - y; sum = t; }
  if (sum != 0.0f) for (x = 0; x < n; x++) v[x] /= sum;
  else             for (x = 0; x < n; x++) v[x] = 1.0f / (float) n;
}

int bo_fchoose(BO_RNG *r, const float *p, int n)
{
  double roll = bo_random(r), sum = 0.0;
  int    i;
  for (i = 0; i < n; i++) { sum += p[i]; if (roll < sum) return i; }
  do { i = (int)(bo_random(r) * n); } while (p[i] == 0.0f);
  return i;
}

/* ---- select_*_fs (stotrace_fs.c:150-365) ---- */
static int select_m(BO_RNG *r, const BO_FS_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  static const int state[4] = { BO_ST_B, BO_ST_M, BO_ST_I, BO_ST_D };
  float path[4];
  path[0] = XM(ox, i, BO_XC_B) * TF(BO_T_BM, k - 1);
  path[1] = (k > 1) ? FC(ox, i, k - 1, BO_FS_M) * TF(BO_T_MM, k - 1) : 0.0f;     /* node 0 is the zero shifted in by rightshiftz */
  path[2] = (k > 1) ? FC(ox, i, k - 1, BO_FS_I) * TF(BO_T_IM, k - 1) : 0.0f;
  path[3] = (k > 1) ? FC(ox, i, k - 1, BO_FS_D) * TF(BO_T_DM, k - 1) : 0.0f;
  bo_fnorm(path, 4);
  return state[bo_fchoose(r, path, 4)];
}

static int select_d(BO_RNG *r, const BO_FS_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  static const int state[2] = { BO_ST_M, BO_ST_D };
  float path[2];
  path[0] = (k > 1) ? FC(ox, i, k - 1, BO_FS_M) * TF(BO_T_MD, k - 1) : 0.0f;
  path[1] = (k > 1) ? FC(ox, i, k - 1, BO_FS_D) * TF(BO_T_DD, k - 1) : 0.0f;
  bo_fnorm(path, 2);
  return state[bo_fchoose(r, path, 2)];
}

static int select_i(BO_RNG *r, const BO_FS_OPROFILE *om, const BO_MX *ox, int i, int k)
{
  static const int state[2] = { BO_ST_M, BO_ST_I };
  float path[2];
  if (i < 3) return -1;                     /* the reference would read before the matrix */
  path[0] = FC(ox, i - 3, k, BO_FS_M) * TF(BO_T_MI, k);
  path[1] = FC(ox, i - 3, k, BO_FS_I) * TF(BO_T_II, k);
  bo_fnorm(path, 2);
  return state[bo_fchoose(r, path, 2)];
}

static int select_cj(BO_RNG *r, const BO_FS_OPROFILE *om, const BO_MX *ox, int i, int cell, float loop, float e_odds, int self)
{
  float path[4], s2, s1, s0;
  if (i < 4) return BO_ST_E;
  s2 = XM(ox, i - 2, BO_XC_SCALE); s1 = XM(ox, i - 1, BO_XC_SCALE); s0 = XM(ox, i, BO_XC_SCALE);
  path[0] = XM(ox, i - 3, cell) * loop;
  path[1] = XM(ox, i - 2, cell) * loop * s2;
  path[2] = XM(ox, i - 1, cell) * loop * s2 * s1;
  path[3] = XM(ox, i, BO_XC_E) * e_odds * s2 * s1 * s0;
  bo_fnorm(path, 4);
  return (bo_fchoose(r, path, 4) < 3) ? self : BO_ST_E;
}

static int select_e(BO_RNG *r, const BO_MX *ox, int i, int *ret_k)
{
  const int M = ox->M;
  const int Q = (((M - 1) / 4) + 1 > 2) ? ((M - 1) / 4) + 1 : 2;          /* p7O_NQF, impl_sse.h:26 */
  double sum = 0.0, roll = bo_random(r), norm = 1.0 / XM(ox, i, BO_XC_E);
  const float xEv = (float) norm;
  int q, z, k, pass;
  for (pass = 0; pass < 1000; pass++)
    for (q = 0; q < Q; q++) {
      for (z = 0; z < 4; z++) { k = z * Q + q + 1; sum += (k <= M) ? FC(ox, i, k, BO_FS_M) * xEv : 0.0f; if (roll < sum) { *ret_k = k; return BO_ST_M; } }
      for (z = 0; z < 4; z++) { k = z * Q + q + 1; sum += (k <= M) ? FC(ox, i, k, BO_FS_D) * xEv : 0.0f; if (roll < sum) { *ret_k = k; return BO_ST_D; } }
    }
  return -1;
}

static int select_b(BO_RNG *r, const BO_FS_OPROFILE *om, const BO_MX *ox, int i)
{
  static const int state[2] = { BO_ST_N, BO_ST_J };
  float path[2];
  path[0] = XM(ox, i, BO_XC_N) * om->xf[BO_X_N][BO_O_MOVE];
  path[1] = XM(ox, i, BO_XC_J) * om->xf[BO_X_J][BO_O_MOVE];
  bo_fnorm(path, 2);
  return state[bo_fchoose(r, path, 2)];
}

static int select_codon_len(BO_RNG *r, const BO_MX *ox, int i, int k)
{
  float path[5];
  int   c;
  for (c = 0; c < 5; c++) path[c] = FC(ox, i, k, BO_FS_M + 1 + c);
  bo_fnorm(path, 5);
  return bo_fchoose(r, path, 5) + 1;
}

/* p7_StochasticTrace_Frameshift (stotrace_fs.c:72-128) */
int bo_StochasticTrace_Frameshift(BO_RNG *rng, int L, const BO_FS_OPROFILE *om, const BO_MX *ox, BO_TRACE *tr)
{
  int i = L, k = 0, c = 0, s0, s1;
  bo_trace_append(tr, BO_ST_T, k, i, c, 0.0f);
  bo_trace_append(tr, BO_ST_C, k, i, c, 0.0f);
  s0 = BO_ST_C;
  while (s0 != BO_ST_S) {
    switch (s0) {
    case BO_ST_M: s1 = select_m(rng, om, ox, i, k); k--;    break;
    case BO_ST_D: s1 = select_d(rng, om, ox, i, k); k--;    break;
    case BO_ST_I: s1 = select_i(rng, om, ox, i, k); i -= 3; break;
    case BO_ST_N: s1 = (i == 0) ? BO_ST_S : BO_ST_N;        break;
    case BO_ST_C: s1 = select_cj(rng, om, ox, i, BO_XC_C, om->xf[BO_X_C][BO_O_LOOP], om->xf[BO_X_E][BO_O_MOVE], BO_ST_C); break;
    case BO_ST_J: s1 = select_cj(rng, om, ox, i, BO_XC_J, om->xf[BO_X_J][BO_O_LOOP], om->xf[BO_X_E][BO_O_LOOP], BO_ST_J); break;
    case BO_ST_E: s1 = select_e(rng, ox, i, &k);            break;
    case BO_ST_B: s1 = select_b(rng, om, ox, i);            break;
    default: return BO_EINVAL;
    }
    if (s1 == -1) return BO_EINVAL;
    if (s1 == BO_ST_M) { c = select_codon_len(rng, ox, i, k); if (i - c < 0) s1 = BO_ST_B; }
    else c = 0;
    bo_trace_append(tr, (char) s1, k, i, c, 0.0f);
    if ((s1 == BO_ST_N || s1 == BO_ST_C || s1 == BO_ST_J) && s1 == s0) i--;
    s0 = s1;
    i -= c;
    if (i < 0) return BO_EINVAL;           /* the reference would index before the matrix here */
  }
  tr->M = om->M; tr->L = L;
  bo_trace_reverse(tr);
  return BO_OK;
}

/* p7_trace_fs_Index (p7_trace.c:2645-2680): domain d covers sqfrom..sqto on the sequence and hmmfrom..hmmto on the model */
int bo_trace_fs_Index(const BO_TRACE *tr, BO_SEGMENT *seg, int max_seg)
{
  int z, nd = 0;
  for (z = 0; z < tr->N; z++)
    switch (tr->st[z]) {
    case BO_ST_B:
      if (nd >= max_seg) return nd;
      seg[nd].i = 0; seg[nd].k = 0; seg[nd].j = 0; seg[nd].m = 0;
      break;
    case BO_ST_M:
      if (seg[nd].i == 0) seg[nd].i = tr->i[z] - tr->c[z] + 1;
      if (seg[nd].k == 0) seg[nd].k = tr->k[z];
      seg[nd].j = tr->i[z]; seg[nd].m = tr->k[z];
      break;
    case BO_ST_E: nd++; break;
    default: break;
    }
  return nd;
}

/* link_spsamples_fs (p7_spensemble.c:226-256) */
static int link_protein = 0;    /* link_spsamples (p7_spensemble.c:191-218) instead of link_spsamples_fs: residue, not nucleotide, diagonals */
static int link_fs(const BO_SEGMENT *h1, const BO_SEGMENT *h2, float min_overlap, int of_smaller, int max_diagdiff)
{
  int nov, n, d1, d2;
#define MIN_(a,b) ((a) < (b) ? (a) : (b))
#define MAX_(a,b) ((a) > (b) ? (a) : (b))
  nov = MIN_(h1->j, h2->j) - MAX_(h1->i, h2->i) + 1;
  n   = of_smaller ? MIN_(h1->j - h1->i + 1, h2->j - h2->i + 1) : MAX_(h1->j - h1->i + 1, h2->j - h2->i + 1);
  if ((float) nov / (float) n < min_overlap) return 0;
  nov = MIN_(h1->m, h2->m) - MAX_(h1->k, h2->k);
  n   = of_smaller ? MIN_(h1->m - h1->k + 1, h2->m - h2->k + 1) : MAX_(h1->m - h1->k + 1, h2->m - h2->k + 1);
  if ((float) nov / (float) n < min_overlap) return 0;
  if (link_protein) {
    d1 = h1->i - h1->k; d2 = h2->i - h2->k; if (abs(d1 - d2) <= max_diagdiff) return 1;
    d1 = h1->j - h1->m; d2 = h2->j - h2->m; if (abs(d1 - d2) <= max_diagdiff) return 1;
    return 0;
  }
  d1 = (h1->i / 3) - h1->k; d2 = (h2->i / 3) - h2->k; if (abs(d1 - d2) <= max_diagdiff) return 1;
  d1 = (h1->j / 3) - h1->m; d2 = (h2->j / 3) - h2->m; if (abs(d1 - d2) <= max_diagdiff) return 1;
  return 0;
}

static int by_start(const void *a, const void *b)
{
  const BO_SEGMENT *x = a, *y = b;
  return (x->i < y->i) ? -1 : (x->i > y->i) ? 1 : 0;
}

/* p7_spensemble_fs_Cluster (p7_spensemble.c:498-640) with Easel's esl_cluster_SingleLinkage, then the removal of
 * dominated clusters of region_trace_ensemble_frameshift (p7_domaindef.c:923-952).  sp[0..n): sampled segments with
 * idx = trace number, in sampling order; out[]: consensus segments ordered by start.  Returns their number. */
int bo_spensemble_fs_Cluster(const BO_SEGMENT *sp, int n, int nsamples, float min_overlap, int of_smaller, int max_diagdiff,
                             float min_posterior, float min_endpointp, BO_SEGMENT *out, int max_out)
{
  int *a = malloc(sizeof(int) * (size_t)(2 * n + 2)), *b = a + n + 1, *asg = malloc(sizeof(int) * (size_t)(n + 1));
  int na = n, nb = 0, nc = 0, v, w, i, c, h, nsig = 0, d, d2;
  for (v = 0; v < n; v++) a[v] = n - v - 1;
  while (na > 0) {
    v = a[na - 1]; na--;
    b[nb++] = v;
    while (nb > 0) {
      v = b[nb - 1]; nb--;
      asg[v] = nc;
      for (i = na - 1; i >= 0; i--)
        if (link_fs(&sp[v], &sp[a[i]], min_overlap, of_smaller, max_diagdiff)) { w = a[i]; a[i] = a[na - 1]; na--; b[nb++] = w; }
    }
    nc++;
  }
  for (c = 0; c < nc && nsig < max_out; c++) {
    int ninc = 0, last = -1, imin = -1, imax = 0, jmin = 0, jmax = 0, kmin = 0, kmax = 0, mmin = 0, mmax = 0, width, thr, *epc, bi, bj, bk, bm, z, best;
    for (h = 0; h < n; h++) if (asg[h] == c) { if (sp[h].idx != last) ninc++; last = sp[h].idx; }
    if ((float) ninc / (float) nsamples < min_posterior) continue;
    for (h = 0; h < n; h++) if (asg[h] == c) {
      if (imin == -1) { imin = imax = sp[h].i; jmin = jmax = sp[h].j; kmin = kmax = sp[h].k; mmin = mmax = sp[h].m; }
      else {
        imin = MIN_(imin, sp[h].i); imax = MAX_(imax, sp[h].i); jmin = MIN_(jmin, sp[h].j); jmax = MAX_(jmax, sp[h].j);
        kmin = MIN_(kmin, sp[h].k); kmax = MAX_(kmax, sp[h].k); mmin = MIN_(mmin, sp[h].m); mmax = MAX_(mmax, sp[h].m);
      }
    }
    width = MAX_(MAX_(imax - imin + 1, jmax - jmin + 1), MAX_(kmax - kmin + 1, mmax - mmin + 1));
    epc = calloc((size_t) width, sizeof(int));
    thr = (int) ceilf((float) ninc * min_endpointp);
#define ARGMAX_(len) do { best = 0; for (z = 1; z < (len); z++) if (epc[z] > epc[best]) best = z; } while (0)
    memset(epc, 0, sizeof(int) * (size_t) width);
    for (h = 0; h < n; h++) if (asg[h] == c) epc[sp[h].i - imin]++;
    for (bi = imin; bi <= imax; bi++) if (epc[bi - imin] >= thr) break;
    if (bi > imax) { ARGMAX_(imax - imin + 1); bi = imin + best; }
    memset(epc, 0, sizeof(int) * (size_t) width);
    for (h = 0; h < n; h++) if (asg[h] == c) epc[sp[h].k - kmin]++;
    for (bk = kmin; bk <= kmax; bk++) if (epc[bk - kmin] >= thr) break;
    if (bk > kmax) { ARGMAX_(kmax - kmin + 1); bk = kmin + best; }
    memset(epc, 0, sizeof(int) * (size_t) width);
    for (h = 0; h < n; h++) if (asg[h] == c) epc[sp[h].j - jmin]++;
    for (bj = jmax; bj >= jmin; bj--) if (epc[bj - jmin] >= thr) break;
    if (bj < jmin) { ARGMAX_(jmax - jmin + 1); bj = jmin + best; }
    memset(epc, 0, sizeof(int) * (size_t) width);
    for (h = 0; h < n; h++) if (asg[h] == c) epc[sp[h].m - mmin]++;
    for (bm = mmax; bm >= mmin; bm--) if (epc[bm - mmin] >= thr) break;
    if (bm < mmin) { ARGMAX_(mmax - mmin + 1); bm = mmin + best; }
    free(epc);
    if (bi > bj || bk > bm) continue;
    out[nsig].i = bi; out[nsig].j = bj; out[nsig].k = bk; out[nsig].m = bm; out[nsig].idx = c;
    out[nsig].prob = (float) ninc / (float) nsamples;
    nsig++;
  }
  qsort(out, (size_t) nsig, sizeof(BO_SEGMENT), by_start);
  /* dominated clusters (p7_domaindef.c:923-952); the flags reuse the assignment array as the reference does */
  for (d = 0; d < nsig; d++) asg[d] = 0;
  for (d = 0; d < nsig; d++)
    for (d2 = d + 1; d2 < nsig; d2++) {
      int nov = MIN_(out[d].j, out[d2].j) - MAX_(out[d].i, out[d2].i) + 1, nn;
      if (nov == 0) break;
      nn = MIN_(out[d].j - out[d].i + 1, out[d2].j - out[d2].i + 1);
      if ((float) nov / (float) nn >= 0.8f) { if (out[d].prob > out[d2].prob) asg[d2] = 1; else asg[d] = 1; }
    }
  for (d = 0, d2 = 0; d2 < nsig; d2++) { if (asg[d2]) continue; if (d != d2) out[d] = out[d2]; d++; }
  free(a); free(asg);
  return d;
}

/* region_trace_ensemble_frameshift (p7_domaindef.c:892-954) on a filled multihit Forward matrix of region ireg..jreg:
 * writes the sampled segments (window coordinates) to samples[] if given, the consensus envelopes to out[]. */
int bo_region_trace_ensemble_frameshift(const BO_FS_OPROFILE *om, const BO_MX *fwd, int ireg, int jreg, uint32_t seed, int nsamples,
                                        BO_SEGMENT *samples, int max_samples, int *ret_nsamples, BO_SEGMENT *out, int max_out)
{
  const int Lr = jreg - ireg + 1;
  BO_RNG rng;
  BO_TRACE *tr = bo_trace_create();
  BO_SEGMENT *sp = malloc(sizeof(BO_SEGMENT) * (size_t) nsamples * 64), seg[64];
  int t, d, n = 0, nd, nc;
  bo_rng_init(&rng, seed);
  for (t = 0; t < nsamples; t++) {
    if (bo_StochasticTrace_Frameshift(&rng, Lr, om, fwd, tr) != BO_OK) { bo_trace_destroy(tr); free(sp); return -1; }
    nd = bo_trace_fs_Index(tr, seg, 64);
    for (d = 0; d < nd; d++) {
      sp[n].idx = t; sp[n].i = seg[d].i + ireg - 1; sp[n].j = seg[d].j + ireg - 1; sp[n].k = seg[d].k; sp[n].m = seg[d].m; sp[n].prob = 0.0f;
      n++;
    }
    bo_trace_reuse(tr);
  }
  if (samples) for (d = 0; d < n && d < max_samples; d++) samples[d] = sp[d];
  if (ret_nsamples) *ret_nsamples = n;
  nc = bo_spensemble_fs_Cluster(sp, n, nsamples, 0.8f, 1, 4, 0.25f, 0.02f, out, max_out);
  bo_trace_destroy(tr); free(sp);
  return nc;
}

/* p7_spensemble_Cluster (p7_spensemble.c:300-440): the same procedure with the protein link rule.  Not re-entrant (tests call it
 * from one thread). */
int bo_spensemble_Cluster(const BO_SEGMENT *sp, int n, int nsamples, float min_overlap, int of_smaller, int max_diagdiff,
                          float min_posterior, float min_endpointp, BO_SEGMENT *out, int max_out)
{
  int nc;
  link_protein = 1;
  nc = bo_spensemble_fs_Cluster(sp, n, nsamples, min_overlap, of_smaller, max_diagdiff, min_posterior, min_endpointp, out, max_out);
  link_protein = 0;
  return nc;
}
