/* filters.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 *
 * The integer filter score systems and filters of the protein (ORF) stage, un-striped:
 *   mf_conversion / vf_conversion / biased_byteify / wordify     src/impl_sse/p7_oprofile.c:667-921
 *   p7_oprofile_ReconfigLength                                    src/impl_sse/p7_oprofile.c:1261-1326
 *   p7_oprofile_GetSSVEmissionScoreArray (P7_SCOREDATA.ssv_scores) :1483-1526
 *   p7_SSVFilter, get_xE                                          src/impl_sse/ssvfilter.c:831-925 (semantics :14-210)
 *   p7_MSVFilter                                                  src/impl_sse/msvfilter.c:74-208
 *   p7_SSVFilter_BATH                                             src/impl_sse/msvfilter.c:250-427
 *   p7_ViterbiFilter, p7_ViterbiFilter_BATH                       src/impl_sse/vitfilter.c:83-248, :286-465
 * Striping is a memory layout; what is kept from it is what leaks into results: the saturating 8/16-bit
 * arithmetic, the lazy-F decision (whether D->D paths are evaluated on a row), and the scan ORDER in which
 * the window finders pick a model position among equal cells (k = q + Q z + 1, q outer; Q depends on the
 * lanes per vector of the CPU build: 16/8 for SSE, 32/16 for AVX2 -- a parameter here).
 * Phantom cells k > M of the striped vectors are not modelled: their costs are saturated so they can only
 * win a row maximum if every real cell is below -12768 (words) / equal to 0 (bytes), which real profiles never reach.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "bath_oracle.h"

#define LOG2 0.69314718055994529

static inline uint8_t u8_adds(uint8_t a, uint8_t b) { int s = (int) a + b; return (uint8_t)(s > 255 ? 255 : s); }
static inline uint8_t u8_subs(uint8_t a, uint8_t b) { int s = (int) a - b; return (uint8_t)(s < 0 ? 0 : s); }
static inline uint8_t u8_max(uint8_t a, uint8_t b)  { return a > b ? a : b; }
static inline int16_t w_adds(int16_t a, int16_t b)  { int s = (int) a + b; return (int16_t)(s > 32767 ? 32767 : (s < -32768 ? -32768 : s)); }
static inline int16_t w_max(int16_t a, int16_t b)   { return a > b ? a : b; }

/* p7_oprofile.c:667-705 */
static uint8_t unbiased_byteify(const BO_OPROFILE *om, float sc)
{
  sc = -1.0f * roundf(om->scale_b * sc);
  return (sc > 255.) ? 255 : (uint8_t) sc;
}
static uint8_t biased_byteify(const BO_OPROFILE *om, float sc)
{
  uint8_t b;
  sc = -1.0f * roundf(om->scale_b * sc);
  b  = (sc > 255 - om->bias_b) ? 255 : (uint8_t) sc + om->bias_b;
  return b;
}
static int16_t wordify(const BO_OPROFILE *om, float sc)
{
  sc = roundf(om->scale_w * sc);
  if      (sc >=  32767.0) return  32767;
  else if (sc <= -32768.0) return -32768;
  else return (int16_t) sc;
}

#define MSC(gm,k,x)  ((gm)->rsc[((size_t)(x) * ((gm)->M + 1) + (k)) * 2 + 0])
#define ISC(gm,k,x)  ((gm)->rsc[((size_t)(x) * ((gm)->M + 1) + (k)) * 2 + 1])
#define TSC(gm,k,t)  ((gm)->tsc[(size_t)(k) * BO_P_NTRANS + (t)])

void bo_oprofile_destroy(BO_OPROFILE *om)
{
  if (!om) return;
  free(om->rbv); free(om->rwv); free(om->twv); free(om->rfv); free(om->tfv); free(om);
}

/* p7_oprofile_Convert: mf_conversion (:773-812), vf_conversion (:826-921), fb_conversion (:925-985) */
BO_OPROFILE *bo_oprofile_convert(const BO_PROFILE *gm)
{
  BO_OPROFILE *om = calloc(1, sizeof(BO_OPROFILE));
  int   M = gm->M, x, k, z, t;
  float max = 0.0;
  static const int gmap[8] = { BO_P_BM, BO_P_MM, BO_P_IM, BO_P_DM, BO_P_MD, BO_P_MI, BO_P_II, BO_P_DD };

  if (!om) return NULL;
  om->M = M; om->L = gm->L; om->mode = gm->mode; om->nj = gm->nj; om->max_length = gm->max_length;
  for (z = 0; z < 8; z++)    om->evparam[z] = gm->evparam[z];
  for (z = 0; z < BO_K; z++) om->compo[z]   = gm->compo[z];
  om->rbv = malloc((size_t) BO_KP * (M + 1));
  om->rwv = malloc(sizeof(int16_t) * (size_t) BO_KP * (M + 1));
  om->twv = malloc(sizeof(int16_t) * (size_t) 8 * (M + 1));
  om->rfv = malloc(sizeof(float) * (size_t) BO_KP * (M + 1));
  om->tfv = malloc(sizeof(float) * (size_t) 8 * (M + 1));
  if (!om->rbv || !om->rwv || !om->twv || !om->rfv || !om->tfv) { bo_oprofile_destroy(om); return NULL; }

  /* ---- bytes */
  for (x = 0; x < BO_K; x++)
    for (k = 0; k <= M; k++) {       /* esl_vec_FMax over match AND insert scores, position 0 included */
      if (MSC(gm, k, x) > max) max = MSC(gm, k, x);
      if (ISC(gm, k, x) > max) max = ISC(gm, k, x);
    }
  om->scale_b = 3.0 / LOG2;
  om->base_b  = 190;
  om->bias_b  = unbiased_byteify(om, -1.0 * max);
  for (x = 0; x < BO_KP; x++) {
    om->rbv[(size_t) x * (M + 1)] = 255;
    for (k = 1; k <= M; k++) om->rbv[(size_t) x * (M + 1) + k] = biased_byteify(om, MSC(gm, k, x));
  }
  om->tbm_b = unbiased_byteify(om, logf(2.0f / ((float) M * (float) (M + 1))));
  om->tec_b = unbiased_byteify(om, logf(0.5f));
  om->tjb_b = unbiased_byteify(om, logf(3.0f / (float) (gm->L + 3)));

  /* ---- words */
  om->scale_w = 500.0 / LOG2;
  om->base_w  = 12000;
  for (x = 0; x < BO_KP; x++) {
    om->rwv[(size_t) x * (M + 1)] = -32768;
    for (k = 1; k <= M; k++) om->rwv[(size_t) x * (M + 1) + k] = wordify(om, MSC(gm, k, x));
  }
  for (t = 0; t < 8; t++) {
    int16_t maxval = (t == BO_T_II) ? -1 : 0;      /* "do not allow an II transition cost of 0" (:872-877) */
    for (k = 0; k <= M; k++) {
      int16_t val = (k < M) ? wordify(om, TSC(gm, k, gmap[t])) : -32768;
      if (t != BO_T_DD && val > maxval) val = maxval;
      om->twv[(size_t) t * (M + 1) + k] = val;
    }
  }
  /* the straight (un-rotated) vectors never hold node 0; DD starts at node 1 */
  om->twv[(size_t) BO_T_MD * (M + 1)] = om->twv[(size_t) BO_T_MI * (M + 1)] = -32768;
  om->twv[(size_t) BO_T_II * (M + 1)] = om->twv[(size_t) BO_T_DD * (M + 1)] = -32768;
  om->xw[BO_X_E][BO_O_LOOP] = wordify(om, gm->xsc[BO_X_E][BO_X_LOOP]);
  om->xw[BO_X_E][BO_O_MOVE] = wordify(om, gm->xsc[BO_X_E][BO_X_MOVE]);
  om->xw[BO_X_N][BO_O_MOVE] = wordify(om, gm->xsc[BO_X_N][BO_X_MOVE]);
  om->xw[BO_X_N][BO_O_LOOP] = 0;
  om->xw[BO_X_C][BO_O_MOVE] = wordify(om, gm->xsc[BO_X_C][BO_X_MOVE]);
  om->xw[BO_X_C][BO_O_LOOP] = 0;
  om->xw[BO_X_J][BO_O_MOVE] = wordify(om, gm->xsc[BO_X_J][BO_X_MOVE]);
  om->xw[BO_X_J][BO_O_LOOP] = 0;
  om->ddbound_w = -32768;
  for (k = 2; k < M - 1; k++) {
    int ddtmp = (int) wordify(om, TSC(gm, k, BO_P_DD));
    ddtmp    += (int) wordify(om, TSC(gm, k + 1, BO_P_DM));
    ddtmp    -= (int) wordify(om, TSC(gm, k + 1, BO_P_BM));
    if (ddtmp > om->ddbound_w) om->ddbound_w = (int16_t) ddtmp;
  }

  /* ---- floats (odds ratios), esl_sse_expf for tables, libm expf for specials */
  for (x = 0; x < BO_KP; x++) {
    om->rfv[(size_t) x * (M + 1)] = 0.0f;
    for (k = 1; k <= M; k++) om->rfv[(size_t) x * (M + 1) + k] = bo_cephes_expf(MSC(gm, k, x));
  }
  for (t = 0; t < 8; t++) {
    for (k = 0; k < M; k++) om->tfv[(size_t) t * (M + 1) + k] = bo_cephes_expf(TSC(gm, k, gmap[t]));
    om->tfv[(size_t) t * (M + 1) + M] = 0.0f;
  }
  om->tfv[(size_t) BO_T_MD * (M + 1)] = om->tfv[(size_t) BO_T_MI * (M + 1)] = 0.0f;
  om->tfv[(size_t) BO_T_II * (M + 1)] = om->tfv[(size_t) BO_T_DD * (M + 1)] = 0.0f;
  for (z = 0; z < 4; z++) {
    om->xf[z][BO_O_LOOP] = expf(gm->xsc[z][BO_X_LOOP]);
    om->xf[z][BO_O_MOVE] = expf(gm->xsc[z][BO_X_MOVE]);
  }
  return om;
}

/* p7_oprofile.c:1261-1326 */
void bo_oprofile_ReconfigLength(BO_OPROFILE *om, int L)
{
  float pmove = (2.0f + om->nj) / ((float) L + 2.0f + om->nj);
  float ploop = 1.0f - pmove;
  om->tjb_b = unbiased_byteify(om, logf(3.0f / (float) (L + 3)));
  om->xf[BO_X_N][BO_O_LOOP] = om->xf[BO_X_C][BO_O_LOOP] = om->xf[BO_X_J][BO_O_LOOP] = ploop;
  om->xf[BO_X_N][BO_O_MOVE] = om->xf[BO_X_C][BO_O_MOVE] = om->xf[BO_X_J][BO_O_MOVE] = pmove;
  om->xw[BO_X_N][BO_O_MOVE] = om->xw[BO_X_C][BO_O_MOVE] = om->xw[BO_X_J][BO_O_MOVE] = wordify(om, logf(pmove));
  om->L = L;
}

/* P7_SCOREDATA.ssv_scores (p7_oprofile.c:1507-1526): byte costs [k*Kp + x] */
void bo_oprofile_ssv_scores(const BO_OPROFILE *om, uint8_t *arr)
{
  int k, x, M = om->M;
  for (x = 0; x < BO_KP; x++) arr[x] = 255;
  for (k = 1; k <= M; k++)
    for (x = 0; x < BO_KP; x++) arr[(size_t) k * BO_KP + x] = om->rbv[(size_t) x * (M + 1) + k];
}

/* ------------------------------------------------------------------------ */
/* get_xE (ssvfilter.c:831-874): every diagonal starts at signed -128 and is stepped with a SIGNED saturating
 * subtraction of sbv = rbv - bias (:750-757); the running maximum is taken UNSIGNED.  Returned as the
 * reference's uint8. */
static uint8_t ssv_get_xE(const uint8_t *dsq, int L, const BO_OPROFILE *om)
{
  int      M = om->M, i, k;
  int8_t  *prev = malloc((size_t) M + 2), *cur = malloc((size_t) M + 2), *tmp;
  uint8_t  xE = 128;      /* beginv as unsigned */
  for (k = 0; k <= M; k++) prev[k] = -128;
  for (i = 1; i <= L; i++) {
    const uint8_t *rb = om->rbv + (size_t) dsq[i] * (M + 1);
    cur[0] = -128;
    for (k = 1; k <= M; k++) {
      /* sbv = ((127 + bias) -sat rbv) ^ 127, read as signed (:750-757) */
      int8_t sb = (int8_t)(u8_subs((uint8_t)(om->bias_b + 127), rb[k]) ^ 127);
      int    s  = (int) prev[k-1] - (int) sb;
      if (s > 127) s = 127;
      if (s < -128) s = -128;
      cur[k] = (int8_t) s;
      if ((uint8_t) cur[k] > xE) xE = (uint8_t) cur[k];
      /* a wrapped vector gets |= 0x80 (:368-373); it only changes values that have already crossed into
       * the non-negative range, i.e. after xE has recorded an overflow -- and then xE is all that is read */
    }
    tmp = prev; prev = cur; cur = tmp;
  }
  free(prev); free(cur);
  return xE;
}

/* ssvfilter.c:876-925 */
int bo_SSVFilter(const uint8_t *dsq, int L, const BO_OPROFILE *om, float *ret_sc)
{
  uint16_t xE, xJ;
  if (om->tjb_b + om->tbm_b + om->tec_b + om->bias_b >= 127) return BO_ENORESULT;
  xE = ssv_get_xE(dsq, L, om);
  if (xE >= 255 - om->bias_b) {
    *ret_sc = INFINITY;
    if (om->base_b - om->tjb_b - om->tbm_b < 128) return BO_ENORESULT;
    return BO_ERANGE;
  }
  xE += om->base_b - om->tjb_b - om->tbm_b;
  xE -= 128;
  if (xE >= 255 - om->bias_b) { *ret_sc = INFINITY; return BO_ERANGE; }
  xJ = xE - om->tec_b;
  if (xJ > om->base_b) return BO_ENORESULT;
  *ret_sc  = ((float) (xJ - om->tjb_b) - (float) om->base_b);
  *ret_sc /= om->scale_b;
  *ret_sc -= 3.0;
  return BO_OK;
}

/* msvfilter.c:74-208.  use_ssv = 0 skips the SSV shortcut (:102-104): the property the shortcut rests on --
 * same score and status either way (ssvfilter.c:14-210) -- is what lets a GPU run the J-state recursion only. */
int bo_MSVFilter_opt(const uint8_t *dsq, int L, const BO_OPROFILE *om, int use_ssv, float *ret_sc)
{
  int      M = om->M, i, k, status;
  uint8_t *prev, *cur, *tmp;
  uint8_t  xJ, xB, xE;
  const uint8_t tjbm = (uint8_t)((int8_t) om->tjb_b + (int8_t) om->tbm_b);

  if (use_ssv) {
    status = bo_SSVFilter(dsq, L, om, ret_sc);
    if (status != BO_ENORESULT) return status;
  }

  prev = calloc((size_t) M + 2, 1); cur = calloc((size_t) M + 2, 1);
  xJ = 0;
  xB = u8_subs(om->base_b, tjbm);
  for (i = 1; i <= L; i++) {
    const uint8_t *rb = om->rbv + (size_t) dsq[i] * (M + 1);
    xE = 0;
    cur[0] = 0;
    for (k = 1; k <= M; k++) {
      uint8_t sv = u8_max(prev[k-1], xB);
      sv = u8_adds(sv, om->bias_b);
      sv = u8_subs(sv, rb[k]);
      if (sv > xE) xE = sv;
      cur[k] = sv;
    }
    if (u8_adds(xE, om->bias_b) == 255) { *ret_sc = INFINITY; free(prev); free(cur); return BO_ERANGE; }
    xE = u8_subs(xE, om->tec_b);
    xJ = u8_max(xJ, xE);
    xB = u8_max(om->base_b, xJ);
    xB = u8_subs(xB, tjbm);
    tmp = prev; prev = cur; cur = tmp;
  }
  free(prev); free(cur);
  *ret_sc  = ((float) (xJ - om->tjb_b) - (float) om->base_b);
  *ret_sc /= om->scale_b;
  *ret_sc -= 3.0;
  return BO_OK;
}

int bo_MSVFilter(const uint8_t *dsq, int L, const BO_OPROFILE *om, float *ret_sc) { return bo_MSVFilter_opt(dsq, L, om, 1, ret_sc); }

/* esl_gumbel_invsurv (Easel esl_gumbel.c): mu - log(-log(1-p)) / lambda, with log1p for small p */
double bo_gumbel_invsurv(double p, double mu, double lambda)
{
  double log_part = (p < 5e-9) ? log(p) : log(-1. * log1p(-p));
  return mu - (log_part / lambda);
}
/* esl_gumbel_surv: 1 - exp(-exp(-lambda (x - mu))), small-argument branch */
double bo_gumbel_surv(double x, double mu, double lambda)
{
  double y  = lambda * (x - mu);
  double ey = -exp(-y);
  if (fabs(ey) < 5e-9) return -ey;
  else                 return 1 - exp(ey);
}
/* esl_exp_surv / esl_exp_logsurv (Easel esl_exponential.c) */
double bo_exp_surv(double x, double mu, double lambda)    { return (x < mu) ? 1.0 : exp(-lambda * (x - mu)); }
double bo_exp_logsurv(double x, double mu, double lambda) { return (x < mu) ? 0.0 : -lambda * (x - mu); }

static int windowlist_add(BO_WINDOWLIST *wl, int n, int k, int length, float score, int target_len)
{
  if (wl->count == wl->nalloc) {
    wl->nalloc = wl->nalloc ? wl->nalloc * 2 : 16;
    wl->w = realloc(wl->w, sizeof(BO_WINDOW) * (size_t) wl->nalloc);
    if (!wl->w) return BO_EMEM;
  }
  wl->w[wl->count].n = n; wl->w[wl->count].k = k; wl->w[wl->count].length = length;
  wl->w[wl->count].score = score; wl->w[wl->count].target_len = target_len; wl->w[wl->count].id = 0;
  wl->count++;
  return BO_OK;
}

/* msvfilter.c:250-427.  lanes = bytes per vector of the CPU build (16 SSE, 32 AVX2): fixes the order in
 * which the cell that crossed the threshold is searched (:358-366). */
int bo_SSVFilter_BATH(const uint8_t *dsq, int L, BO_OPROFILE *om, const uint8_t *ssv_scores, float nullsc, double P,
                      int lanes, BO_WINDOWLIST *wl)
{
  float   invP = bo_gumbel_invsurv(P, om->evparam[BO_MMU], om->evparam[BO_MLAMBDA]);
  uint8_t sc_thresh;
  bo_oprofile_ReconfigLength(om, L);      /* p7_oprofile_ReconfigMSVLength (:308); the caller passes NullOne for L */
  sc_thresh = (uint8_t)(int) ceil(((nullsc + (invP * LOG2) + 3.0) * om->scale_b) + om->base_b + om->tec_b + om->tjb_b);
  return bo_SSVFilter_BATH_thresh(dsq, L, om, ssv_scores, sc_thresh, lanes, wl);
}

/* the same with sc_thresh (:313) given and om->tjb_b already set for L */
int bo_SSVFilter_BATH_thresh(const uint8_t *dsq, int L, const BO_OPROFILE *om, const uint8_t *ssv_scores, uint8_t sc_thresh,
                             int lanes, BO_WINDOWLIST *wl)
{
  int      M = om->M, i, k, q, z;
  int      Q = (M - 1) / lanes + 1;
  uint8_t *prev = calloc((size_t) M + 2, 1), *cur = calloc((size_t) M + 2, 1), *tmp;
  uint8_t  xB, xE;
  uint8_t  tjbm;

  if (Q < 2) Q = 2;
  tjbm = (uint8_t)((int8_t) om->tjb_b + (int8_t) om->tbm_b);
  xB = u8_subs(om->base_b, tjbm);

  for (i = 1; i <= L; i++) {
    const uint8_t *rb = om->rbv + (size_t) dsq[i] * (M + 1);
    xE = 0;
    cur[0] = 0;
    for (k = 1; k <= M; k++) {
      uint8_t sv = u8_max(prev[k-1], xB);
      sv = u8_adds(sv, om->bias_b);
      sv = u8_subs(sv, rb[k]);
      if (sv > xE) xE = sv;
      cur[k] = sv;
    }
    tmp = prev; prev = cur; cur = tmp;     /* prev = row i */
    if (u8_adds(xE, (uint8_t)(255 - sc_thresh)) == 255) {     /* xE >= sc_thresh (:343-347) */
      int end = -1, rem_sc = -1, start, target_start, target_end, sc, n, max_end, max_sc, pos_since_max;
      float ret_sc;
      for (q = 0; q < Q; q++)
        for (z = 0; z < lanes; z++) {
          k = q + Q * z + 1;
          if (k <= M && prev[k] >= sc_thresh && (int) prev[k] > rem_sc) { end = k; rem_sc = prev[k]; }
        }
      for (k = 0; k <= M; k++) prev[k] = 0;                   /* dp reset (:366) */
      start = end; target_end = target_start = i; sc = rem_sc;
      while (rem_sc > om->base_b - om->tjb_b - om->tbm_b) {
        rem_sc -= om->bias_b - ssv_scores[start * BO_KP + dsq[target_start]];
        --start; --target_start;
      }
      start++; target_start++;
      k = end + 1; n = target_end + 1; max_end = target_end; max_sc = sc; pos_since_max = 0;
      while (k < M && n <= L) {
        sc += om->bias_b - ssv_scores[k * BO_KP + dsq[n]];
        if (sc >= max_sc) { max_sc = sc; max_end = n; pos_since_max = 0; }
        else { pos_since_max++; if (pos_since_max == 5) break; }
        k++; n++;
      }
      end += (max_end - target_end);
      target_end = max_end;
      ret_sc  = ((float) (max_sc - om->tjb_b) - (float) om->base_b);
      ret_sc /= om->scale_b;
      ret_sc -= 3.0;
      windowlist_add(wl, target_start, end, end - start + 1, ret_sc, L);
      i = target_end;
    }
  }
  free(prev); free(cur);
  return BO_OK;
}

/* ------------------------------------------------------------------------ */
/* One Viterbi-filter row (vitfilter.c:130-231 == :353-454).  Returns xE; updates the row in place. */
typedef struct { int16_t *m, *d, *i; } vf_row;

static int16_t vf_step(const BO_OPROFILE *om, const int16_t *rw, int16_t xB, vf_row *prev, vf_row *cur, int16_t *ret_Dmax)
{
  int M = om->M, k;
  const int16_t *tBM = om->twv + (size_t) BO_T_BM * (M + 1), *tMM = om->twv + (size_t) BO_T_MM * (M + 1);
  const int16_t *tIM = om->twv + (size_t) BO_T_IM * (M + 1), *tDM = om->twv + (size_t) BO_T_DM * (M + 1);
  const int16_t *tMD = om->twv + (size_t) BO_T_MD * (M + 1), *tMI = om->twv + (size_t) BO_T_MI * (M + 1);
  const int16_t *tII = om->twv + (size_t) BO_T_II * (M + 1);
  int16_t xE = -32768, Dmax = -32768;
  cur->m[0] = cur->d[0] = cur->i[0] = -32768;
  cur->d[1] = -32768;
  for (k = 1; k <= M; k++) {
    int16_t sv = w_adds(xB, tBM[k-1]);
    sv = w_max(sv, w_adds(prev->m[k-1], tMM[k-1]));
    sv = w_max(sv, w_adds(prev->i[k-1], tIM[k-1]));
    sv = w_max(sv, w_adds(prev->d[k-1], tDM[k-1]));
    sv = w_adds(sv, rw[k]);
    if (sv > xE) xE = sv;
    cur->m[k] = sv;
    {
      int16_t dcv = w_adds(sv, tMD[k]);       /* D(i,k+1), M->D only */
      if (dcv > Dmax) Dmax = dcv;
      if (k < M) cur->d[k+1] = dcv;
    }
    cur->i[k] = w_max(w_adds(prev->m[k], tMI[k]), w_adds(prev->i[k], tII[k]));
  }
  *ret_Dmax = Dmax;
  return xE;
}

/* the "lazy F" decision and, when taken, the complete D->D closure (vitfilter.c:197-231) */
static void vf_lazy_dd(const BO_OPROFILE *om, int16_t Dmax, int16_t xB, vf_row *cur)
{
  int M = om->M, k;
  const int16_t *tDD = om->twv + (size_t) BO_T_DD * (M + 1);
  if ((int) Dmax + (int) om->ddbound_w > (int) xB)
    for (k = 2; k <= M; k++) cur->d[k] = w_max(cur->d[k], w_adds(cur->d[k-1], tDD[k-1]));
}

static int vf_alloc(int M, vf_row *a, vf_row *b)
{
  vf_row *r[2] = { a, b };
  int z, k;
  for (z = 0; z < 2; z++) {
    r[z]->m = malloc(sizeof(int16_t) * (size_t)(M + 2));
    r[z]->d = malloc(sizeof(int16_t) * (size_t)(M + 2));
    r[z]->i = malloc(sizeof(int16_t) * (size_t)(M + 2));
    if (!r[z]->m || !r[z]->d || !r[z]->i) return BO_EMEM;
    for (k = 0; k <= M + 1; k++) r[z]->m[k] = r[z]->d[k] = r[z]->i[k] = -32768;
  }
  return BO_OK;
}
static void vf_free(vf_row *a, vf_row *b) { free(a->m); free(a->d); free(a->i); free(b->m); free(b->d); free(b->i); }

/* vitfilter.c:83-248 and :286-465 share everything but the window emission; wl == NULL gives p7_ViterbiFilter. */
int bo_ViterbiFilter_BATH(const uint8_t *dsq, int L, const BO_OPROFILE *om, const uint8_t *ssv_scores, float filtersc, double P,
                          int lanes, BO_WINDOWLIST *wl, float *ret_sc)
{
  int16_t sc_thresh = 0;
  int     sc_ext_thresh = 0;
  if (wl) {
    float invP = bo_gumbel_invsurv(P, om->evparam[BO_VMU], om->evparam[BO_VLAMBDA]);
    sc_thresh = (int16_t) ceil(((filtersc + LOG2 * invP + 3.0) * om->scale_w)
                               - (float) om->xw[BO_X_E][BO_O_MOVE] - (float) om->xw[BO_X_C][BO_O_MOVE] + (float) om->base_w);
    invP = bo_gumbel_invsurv(P, om->evparam[BO_MMU], om->evparam[BO_MLAMBDA]);
    sc_ext_thresh = (int) ceil(((filtersc + LOG2 * invP + 3.0) * om->scale_b) + om->base_b + om->tec_b + om->tjb_b);
  }
  return bo_ViterbiFilter_BATH_thresh(dsq, L, om, ssv_scores, sc_thresh, sc_ext_thresh, lanes, wl, ret_sc);
}

/* the same with both thresholds (:315-321) given */
int bo_ViterbiFilter_BATH_thresh(const uint8_t *dsq, int L, const BO_OPROFILE *om, const uint8_t *ssv_scores, int16_t sc_thresh,
                                 int sc_ext_thresh, int lanes, BO_WINDOWLIST *wl, float *ret_sc)
{
  int     M = om->M, i, k, q, z;
  int     Q = (M - 1) / lanes + 1;
  vf_row  ra, rb, *prev = &ra, *cur = &rb, *tmp;
  int16_t xE, xB, xC, xJ, xN, Dmax;
  int     skip_until = 0;

  if (Q < 2) Q = 2;
  if (vf_alloc(M, &ra, &rb) != BO_OK) return BO_EMEM;
  xN = om->base_w;
  xB = xN + om->xw[BO_X_N][BO_O_MOVE];
  xJ = -32768; xC = -32768; xE = -32768;

  for (i = 1; i <= L; i++) {
    const int16_t *rw = om->rwv + (size_t) dsq[i] * (M + 1);
    xE = vf_step(om, rw, xB, prev, cur, &Dmax);
    if (xE >= 32767) { *ret_sc = INFINITY; vf_free(&ra, &rb); return BO_ERANGE; }
    xN = xN + om->xw[BO_X_N][BO_O_LOOP];
    xC = (xC + om->xw[BO_X_C][BO_O_LOOP] > xE + om->xw[BO_X_E][BO_O_MOVE]) ? xC + om->xw[BO_X_C][BO_O_LOOP] : xE + om->xw[BO_X_E][BO_O_MOVE];
    xJ = (xJ + om->xw[BO_X_J][BO_O_LOOP] > xE + om->xw[BO_X_E][BO_O_LOOP]) ? xJ + om->xw[BO_X_J][BO_O_LOOP] : xE + om->xw[BO_X_E][BO_O_LOOP];
    xB = (xJ + om->xw[BO_X_J][BO_O_MOVE] > xN + om->xw[BO_X_N][BO_O_MOVE]) ? xJ + om->xw[BO_X_J][BO_O_MOVE] : xN + om->xw[BO_X_N][BO_O_MOVE];

    if (wl && i > skip_until && xE >= sc_thresh) {
      int k_start = 0, max_k_end, max_i_end, sc_ext, max_sc_ext, pos_since_max, kk, nn;
      for (q = 0; q < Q && k_start == 0; q++)
        for (z = 0; z < lanes; z++) {
          k = q + Q * z + 1;
          if (k <= M && cur->m[k] == xE) { k_start = k; break; }
        }
      max_k_end = k_start; max_i_end = i; sc_ext = sc_ext_thresh; max_sc_ext = sc_ext; pos_since_max = 0;
      kk = k_start + 1; nn = i + 1;
      while (kk <= M && nn <= L) {
        sc_ext += om->bias_b - ssv_scores[kk * BO_KP + dsq[nn]];
        if (sc_ext >= max_sc_ext) { max_sc_ext = sc_ext; max_k_end = kk; max_i_end = nn; pos_since_max = 0; }
        else if (++pos_since_max == 5) break;
        kk++; nn++;
      }
      windowlist_add(wl, i, max_k_end, max_k_end - k_start + 1, 0.0f, L);
      skip_until = max_i_end;
    }
    vf_lazy_dd(om, Dmax, xB, cur);
    tmp = prev; prev = cur; cur = tmp;
  }
  vf_free(&ra, &rb);
  if (xC > -32768) {
    *ret_sc  = (float) xC + (float) om->xw[BO_X_C][BO_O_MOVE] - (float) om->base_w;
    *ret_sc /= om->scale_w;
    *ret_sc -= 3.0;
  } else *ret_sc = -INFINITY;
  return BO_OK;
}

int bo_ViterbiFilter(const uint8_t *dsq, int L, const BO_OPROFILE *om, float *ret_sc)
{
  return bo_ViterbiFilter_BATH(dsq, L, om, NULL, 0.0f, 0.0, 8, NULL, ret_sc);
}

void bo_windowlist_reset(BO_WINDOWLIST *wl) { wl->count = 0; }
void bo_windowlist_free(BO_WINDOWLIST *wl)  { free(wl->w); wl->w = NULL; wl->count = wl->nalloc = 0; }
