/* profile.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 * Null model, protein profile, frameshift profile and its odds-ratio form.
 * Follows src/p7_bg.c:60-100,189-197,356-384; src/p7_hmm.c:1349-1364;
 * src/modelconfig.c:48-196,220-698,722-874; src/p7_profile.c:152-240;
 * src/impl_sse/p7_fs_oprofile.c:222-296,636-750. */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "bath_oracle.h"

/* ---------------------------------------------------------------- bg */

/* src/hmmer.c:163-182 (Swiss-Prot 50.8) */
static const float AMINO_F[BO_K] = {
  0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198, 0.0590092,
  0.0594422, 0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639, 0.0540978, 0.0683364,
  0.0540687, 0.0673417, 0.0114135, 0.0304133 };

/* src/p7_bg.c:52-82 */
BO_BG *bo_bg_create(void)
{
  BO_BG *bg = calloc(1, sizeof(BO_BG));
  int x;
  if (!bg) return NULL;
  for (x = 0; x < BO_K; x++) bg->f[x] = AMINO_F[x];
  bg->p1    = 350. / 351.;
  bg->omega = 1. / 256.;
  return bg;
}
void bo_bg_destroy(BO_BG *bg) { free(bg); }

/* src/p7_bg.c:189-197 */
void bo_bg_SetLength(BO_BG *bg, int L)
{
  bg->p1 = (float) L / (float) (L + 1);
  bg->fh_t[0][0] = bg->p1;
  bg->fh_t[0][1] = 1.0f - bg->p1;
}

/* src/p7_bg.c:356-360 */
float bo_bg_NullOne(const BO_BG *bg, int L)
{
  return (float) L * log(bg->p1) + log(1. - bg->p1);
}

/* src/p7_bg.c:377-384 */
float bo_bg_fs_NullOne(const BO_BG *bg, int aminoL)
{
  float null_per_frame = (float) aminoL * log(bg->p1) + log(1. - bg->p1);
  return null_per_frame + log(3.0);
}

/* ---------------------------------------------------------------- hmm */

/* src/p7_hmm.c:1349-1364 */
void bo_hmm_CalculateOccupancy(const BO_HMM *hmm, float *mocc)
{
  int k;
  mocc[0] = 0.;
  mocc[1] = hmm->t[0 * 7 + BO_H_MI] + hmm->t[0 * 7 + BO_H_MM];
  for (k = 2; k <= hmm->M; k++)
    mocc[k] = mocc[k-1] * (hmm->t[(k-1) * 7 + BO_H_MM] + hmm->t[(k-1) * 7 + BO_H_MI]) +
              (1.0 - mocc[k-1]) * hmm->t[(k-1) * 7 + BO_H_DM];
}

/* shared by both config functions: entry, E, core transitions
 * (src/modelconfig.c:85-136 == :283-332) */
static int config_transitions(const BO_HMM *hmm, float *tsc, float xsc[4][2], float *nj, int mode)
{
  int    M = hmm->M, k;
  float *occ, Z;

  /* p7_profile_Create: node 0 has no transitions (p7_profile.c:205) */
  for (k = 0; k < BO_P_NTRANS; k++) tsc[k] = -INFINITY;

  if (mode == BO_LOCAL || mode == BO_UNILOCAL) {
    occ = malloc(sizeof(float) * (M + 1));
    if (!occ) return BO_EMEM;
    bo_hmm_CalculateOccupancy(hmm, occ);
    Z = 0.;
    for (k = 1; k <= M; k++) Z += occ[k] * (float) (M - k + 1);
    for (k = 1; k <= M; k++) tsc[(k-1) * BO_P_NTRANS + BO_P_BM] = log(occ[k] / Z);
    free(occ);
  } else {
    Z = log(hmm->t[0 * 7 + BO_H_MD]);
    tsc[0 * BO_P_NTRANS + BO_P_BM] = log(1.0 - hmm->t[0 * 7 + BO_H_MD]);
    for (k = 1; k < M; k++) {
      tsc[k * BO_P_NTRANS + BO_P_BM] = Z + log(hmm->t[k * 7 + BO_H_DM]);
      Z += log(hmm->t[k * 7 + BO_H_DD]);
    }
  }

  if (mode == BO_LOCAL || mode == BO_GLOCAL) {
    xsc[BO_X_E][BO_X_MOVE] = -0.69314718055994529;
    xsc[BO_X_E][BO_X_LOOP] = -0.69314718055994529;
    *nj = 1.0f;
  } else {
    xsc[BO_X_E][BO_X_MOVE] = 0.0f;
    xsc[BO_X_E][BO_X_LOOP] = -INFINITY;
    *nj = 0.0f;
  }

  for (k = 1; k < M; k++) {
    float *tp = tsc + k * BO_P_NTRANS;
    tp[BO_P_MM] = log(hmm->t[k * 7 + BO_H_MM]);
    tp[BO_P_MI] = log(hmm->t[k * 7 + BO_H_MI]);
    tp[BO_P_MD] = log(hmm->t[k * 7 + BO_H_MD]);
    tp[BO_P_IM] = log(hmm->t[k * 7 + BO_H_IM]);
    tp[BO_P_II] = log(hmm->t[k * 7 + BO_H_II]);
    tp[BO_P_DM] = log(hmm->t[k * 7 + BO_H_DM]);
    tp[BO_P_DD] = log(hmm->t[k * 7 + BO_H_DD]);
  }
  return BO_OK;
}

/* src/modelconfig.c:722-736 */
void bo_profile_ReconfigLength(BO_PROFILE *gm, int L)
{
  float pmove = (2.0f + gm->nj) / ((float) L + 2.0f + gm->nj);
  float ploop = 1.0f - pmove;
  gm->xsc[BO_X_N][BO_X_LOOP] = gm->xsc[BO_X_C][BO_X_LOOP] = gm->xsc[BO_X_J][BO_X_LOOP] = log(ploop);
  gm->xsc[BO_X_N][BO_X_MOVE] = gm->xsc[BO_X_C][BO_X_MOVE] = gm->xsc[BO_X_J][BO_X_MOVE] = log(pmove);
  gm->L = L;
}

/* src/modelconfig.c:48-196 */
BO_PROFILE *bo_profile_config(const BO_HMM *hmm, const BO_BG *bg, int L, int mode)
{
  BO_PROFILE *gm = calloc(1, sizeof(BO_PROFILE));
  int   M = hmm->M, k, x, z;
  float sc[BO_KP];

  if (!gm) return NULL;
  gm->M = M; gm->max_length = hmm->max_length; gm->mode = mode;
  gm->tsc = malloc(sizeof(float) * (size_t) M * BO_P_NTRANS);
  gm->rsc = malloc(sizeof(float) * (size_t) BO_KP * (M + 1) * 2);
  if (!gm->tsc || !gm->rsc) { bo_profile_destroy(gm); return NULL; }
  for (z = 0; z < 8; z++)    gm->evparam[z] = hmm->evparam[z];
  for (z = 0; z < BO_K; z++) gm->compo[z]   = hmm->compo[z];

  /* p7_profile_Create edge init: node 0 emissions impossible (p7_profile.c:85-95) */
  for (k = 0; k < M * BO_P_NTRANS; k++) gm->tsc[k] = -INFINITY;
  for (x = 0; x < BO_KP; x++) { gm->rsc[(size_t) x * (M+1) * 2 + 0] = -INFINITY; gm->rsc[(size_t) x * (M+1) * 2 + 1] = -INFINITY; }

  if (config_transitions(hmm, gm->tsc, gm->xsc, &gm->nj, mode) != BO_OK) { bo_profile_destroy(gm); return NULL; }

  sc[BO_K]      = -INFINITY;
  sc[BO_KP - 2] = -INFINITY;
  sc[BO_KP - 1] = -INFINITY;
  for (k = 1; k <= M; k++) {
    for (x = 0; x < BO_K; x++)
      sc[x] = log((double) hmm->mat[k * BO_K + x] / bg->f[x]);
    bo_abc_FExpectScVec(sc, bg->f);
    for (x = 0; x < BO_KP; x++)
      gm->rsc[((size_t) x * (M+1) + k) * 2 + 0] = sc[x];
  }
  for (x = 0; x < BO_KP; x++) {
    for (k = 1; k < M; k++) gm->rsc[((size_t) x * (M+1) + k) * 2 + 1] = 0.0f;
    gm->rsc[((size_t) x * (M+1) + M) * 2 + 1] = -INFINITY;
  }
  for (k = 1; k <= M; k++) {
    gm->rsc[((size_t) BO_K      * (M+1) + k) * 2 + 1] = -INFINITY;
    gm->rsc[((size_t)(BO_KP-2)  * (M+1) + k) * 2 + 1] = -INFINITY;
    gm->rsc[((size_t)(BO_KP-1)  * (M+1) + k) * 2 + 1] = -INFINITY;
  }
  gm->L = 0;
  bo_profile_ReconfigLength(gm, L);
  return gm;
}

void bo_profile_destroy(BO_PROFILE *gm)
{
  if (!gm) return;
  free(gm->tsc); free(gm->rsc); free(gm);
}

/* ---------------------------------------------------------------- fs profile */

/* src/modelconfig.c:760-774 */
void bo_fs_ReconfigLength(BO_FS_PROFILE *gm, int L_amino)
{
  float pmove = (2.0f + gm->nj) / ((float) L_amino + 2.0f + gm->nj);
  float ploop = 1.0f - pmove;
  gm->xsc[BO_X_N][BO_X_LOOP] = gm->xsc[BO_X_C][BO_X_LOOP] = gm->xsc[BO_X_J][BO_X_LOOP] = log(ploop);
  gm->xsc[BO_X_N][BO_X_MOVE] = gm->xsc[BO_X_C][BO_X_MOVE] = gm->xsc[BO_X_J][BO_X_MOVE] = log(pmove);
  gm->L = L_amino;
}
/* src/modelconfig.c:824-831 */
void bo_fs_ReconfigMultihit(BO_FS_PROFILE *gm, int L_amino)
{
  gm->xsc[BO_X_E][BO_X_MOVE] = -0.69314718055994529;
  gm->xsc[BO_X_E][BO_X_LOOP] = -0.69314718055994529;
  gm->nj = 1.0f;
  bo_fs_ReconfigLength(gm, L_amino);
}
/* src/modelconfig.c:867-874 */
void bo_fs_ReconfigUnihit(BO_FS_PROFILE *gm, int L_amino)
{
  gm->xsc[BO_X_E][BO_X_MOVE] = 0.0f;
  gm->xsc[BO_X_E][BO_X_LOOP] = -INFINITY;
  gm->nj = 0.0f;
  bo_fs_ReconfigLength(gm, L_amino);
}

#define RSC(c,k)    (gm->rsc[(size_t)(c) * (M+1) + (k)])
#define AMINOSC(k,a) RSC(maxcodons + (a), (k))
#define CODAA(k,c)  (gm->codons[(size_t)(k) * maxcodons + (c)])
#define CODIN(k,c)  (gm->indel_pos[(size_t)(k) * maxcodons + (c)])
#define TRY(cidx, aa, pat) do { int ci_ = (cidx); int a_ = (aa);                  \
    if (AMINOSC(k, a_) > RSC(ci_, k)) { RSC(ci_, k) = AMINOSC(k, a_);             \
      CODAA(k, ci_) = (uint8_t) a_; CODIN(k, ci_) = (pat); } } while (0)

/* src/modelconfig.c:220-698 */
BO_FS_PROFILE *bo_fs_profile_config(const BO_HMM *hmm, const BO_BG *bg, int ct, int codon_lengths, int L_amino, int mode)
{
  BO_FS_PROFILE *gm;
  const uint8_t *basic = bo_gencode_basic(ct);
  int    M = hmm->M;
  int    maxcodons;
  int    k, t, u, v, w, x, z, a, subn, suba, codon, codon_idx;
  float  sc[BO_KP];
  float  one_indel = 0.f, two_indel = 0.f, no_indel = 0.f, stop_codon = 0.f;

  if (!basic) return NULL;
  if (codon_lengths == 5) {
    maxcodons  = BO_MAXCODONS5;
    one_indel  = log(hmm->fsprob);
    two_indel  = log(hmm->fsprob / 2.);
    stop_codon = log(hmm->fsprob);
    no_indel   = log(1. - hmm->fsprob * 4.);
  } else if (codon_lengths == 3) {
    maxcodons  = BO_MAXCODONS3;
    one_indel  = log(hmm->fsprob);
    stop_codon = log(hmm->fsprob);
    no_indel   = log(1. - hmm->fsprob * 3.);
  } else return NULL;

  gm = calloc(1, sizeof(BO_FS_PROFILE));
  if (!gm) return NULL;
  gm->M = M; gm->max_length = hmm->max_length; gm->mode = mode;
  gm->codon_lengths = codon_lengths; gm->maxcodons = maxcodons; gm->fsprob = hmm->fsprob;
  for (z = 0; z < 8; z++) gm->evparam[z] = hmm->evparam[z];
  gm->tsc       = malloc(sizeof(float) * (size_t) M * BO_P_NTRANS);
  gm->rsc       = malloc(sizeof(float) * (size_t)(maxcodons + BO_KP) * (M + 1));
  gm->codons    = calloc((size_t)(M + 1) * (maxcodons + 1), 1);
  gm->indel_pos = calloc((size_t)(M + 1) * (maxcodons + 1), 1);
  if (!gm->tsc || !gm->rsc || !gm->codons || !gm->indel_pos) { bo_fs_profile_destroy(gm); return NULL; }

  for (k = 0; k < M * BO_P_NTRANS; k++) gm->tsc[k] = -INFINITY;
  if (config_transitions(hmm, gm->tsc, gm->xsc, &gm->nj, mode) != BO_OK) { bo_fs_profile_destroy(gm); return NULL; }

  sc[BO_K]      = -INFINITY;
  sc[BO_KP - 2] = -INFINITY;
  sc[BO_KP - 1] = -INFINITY;
  for (x = 0; x < maxcodons + BO_KP; x++)
    for (k = 0; k <= M; k++) RSC(x, k) = -INFINITY;

  for (k = 1; k <= M; k++) {
    for (x = 0; x < BO_K; x++)
      sc[x] = log((double) hmm->mat[k * BO_K + x] / bg->f[x]);
    bo_abc_FExpectScVec(sc, bg->f);
    for (x = 0; x < BO_KP; x++) RSC(maxcodons + x, k) = sc[x];
  }

  if (codon_lengths == 5) {
    for (k = 1; k <= M; k++) {
      for (x = 0; x < 4; x++)
        for (w = 0; w < 4; w++)
          for (v = 0; v < 4; v++) {
            codon = 16 * v + 4 * w + x;
            a = basic[codon];
            TRY(BO_CODON1_FS5(x),    a, BO___X);
            TRY(BO_CODON1_FS5(v),    a, BO_X__);
            TRY(BO_CODON2_FS5(w, x), a, BO__XX);
            TRY(BO_CODON2_FS5(v, x), a, BO_X_X);
            TRY(BO_CODON2_FS5(v, w), a, BO_XX_);
            codon_idx = BO_CODON3_FS5(v, w, x);
            if (a == BO_KP - 2) {
              for (subn = 0; subn < 4; subn++) {
                suba = basic[16 * subn + 4 * w + x]; TRY(codon_idx, suba, BO_xXX);
                suba = basic[16 * v + 4 * subn + x]; TRY(codon_idx, suba, BO_XxX);
                suba = basic[16 * v + 4 * w + subn]; TRY(codon_idx, suba, BO_XXx);
              }
            } else {
              RSC(codon_idx, k) = AMINOSC(k, a);
              CODAA(k, codon_idx) = (uint8_t) a; CODIN(k, codon_idx) = BO_XXX;
            }
            for (u = 0; u < 4; u++) {
              codon_idx = BO_CODON4_FS5(u, v, w, x);
              a = basic[16 * u + 4 * v + x]; TRY(codon_idx, a, BO_XXxX);
              a = basic[16 * u + 4 * w + x]; TRY(codon_idx, a, BO_XxXX);
              a = basic[16 * v + 4 * w + x]; TRY(codon_idx, a, BO_xXXX);
              for (t = 0; t < 4; t++) {
                codon_idx = BO_CODON5_FS5(t, u, v, w, x);
                a = basic[16 * t + 4 * u + x]; TRY(codon_idx, a, BO_XXxxX);
                a = basic[16 * t + 4 * w + x]; TRY(codon_idx, a, BO_XxxXX);
                a = basic[16 * v + 4 * w + x]; TRY(codon_idx, a, BO_xxXXX);
              }
            }
          }
      /* indel costs (modelconfig.c:497-519) */
      for (x = 0; x < 4; x++) {
        RSC(BO_CODON1_FS5(x), k) += two_indel;
        for (w = 0; w < 4; w++) {
          RSC(BO_CODON2_FS5(w, x), k) += one_indel;
          for (v = 0; v < 4; v++) {
            a = basic[16 * v + 4 * w + x];
            RSC(BO_CODON3_FS5(v, w, x), k) += (a == BO_KP - 2) ? stop_codon : no_indel;
            for (u = 0; u < 4; u++) {
              RSC(BO_CODON4_FS5(u, v, w, x), k) += one_indel;
              for (t = 0; t < 4; t++)
                RSC(BO_CODON5_FS5(t, u, v, w, x), k) += two_indel;
            }
          }
        }
      }
      a = BO_KP - 3;
      RSC(BO_DEGEN5_C,   k) = AMINOSC(k, a) + no_indel;  CODAA(k, BO_DEGEN5_C)   = (uint8_t) a; CODIN(k, BO_DEGEN5_C)   = BO_xxx;
      RSC(BO_DEGEN5_QC1, k) = AMINOSC(k, a) + one_indel; CODAA(k, BO_DEGEN5_QC1) = (uint8_t) a; CODIN(k, BO_DEGEN5_QC1) = BO_xxx;
      RSC(BO_DEGEN5_QC2, k) = AMINOSC(k, a) + two_indel; CODAA(k, BO_DEGEN5_QC2) = (uint8_t) a; CODIN(k, BO_DEGEN5_QC2) = BO_xxx;
    }
  } else {
    for (k = 1; k <= M; k++) {
      for (x = 0; x < 4; x++)
        for (w = 0; w < 4; w++)
          for (v = 0; v < 4; v++) {
            codon = 16 * v + 4 * w + x;
            a = basic[codon];
            TRY(BO_CODON2_FS3(w, x), a, BO__XX);
            TRY(BO_CODON2_FS3(v, x), a, BO_X_X);
            TRY(BO_CODON2_FS3(v, w), a, BO_XX_);
            codon_idx = BO_CODON3_FS3(v, w, x);
            if (a == BO_KP - 2) {
              for (subn = 0; subn < 4; subn++) {
                suba = basic[16 * subn + 4 * w + x]; TRY(codon_idx, suba, BO_xXX);
                suba = basic[16 * v + 4 * subn + x]; TRY(codon_idx, suba, BO_XxX);
                suba = basic[16 * v + 4 * w + subn]; TRY(codon_idx, suba, BO_XXx);
              }
            } else {
              RSC(codon_idx, k) = AMINOSC(k, a);
              CODAA(k, codon_idx) = (uint8_t) a; CODIN(k, codon_idx) = BO_XXX;
            }
            for (u = 0; u < 4; u++) {
              codon_idx = BO_CODON4_FS3(u, v, w, x);
              a = basic[16 * u + 4 * v + x]; TRY(codon_idx, a, BO_XXxX);
              a = basic[16 * u + 4 * w + x]; TRY(codon_idx, a, BO_XxXX);
              a = basic[16 * v + 4 * w + x]; TRY(codon_idx, a, BO_xXXX);
            }
          }
      for (x = 0; x < 4; x++)
        for (w = 0; w < 4; w++) {
          RSC(BO_CODON2_FS3(w, x), k) += one_indel;
          for (v = 0; v < 4; v++) {
            a = basic[16 * v + 4 * w + x];
            RSC(BO_CODON3_FS3(v, w, x), k) += (a == BO_KP - 2) ? stop_codon : no_indel;
            for (u = 0; u < 4; u++)
              RSC(BO_CODON4_FS3(u, v, w, x), k) += one_indel;
          }
        }
      a = BO_KP - 3;
      RSC(BO_DEGEN3_C,   k) = AMINOSC(k, a) + no_indel;  CODAA(k, BO_DEGEN3_C)   = (uint8_t) a; CODIN(k, BO_DEGEN3_C)   = BO_xxx;
      RSC(BO_DEGEN3_QC1, k) = AMINOSC(k, a) + one_indel; CODAA(k, BO_DEGEN3_QC1) = (uint8_t) a; CODIN(k, BO_DEGEN3_QC1) = BO_xxx;
    }
  }

  gm->L = 0;
  bo_fs_ReconfigLength(gm, L_amino);
  return gm;
}
#undef TRY

void bo_fs_profile_destroy(BO_FS_PROFILE *gm)
{
  if (!gm) return;
  free(gm->tsc); free(gm->rsc); free(gm->codons); free(gm->indel_pos); free(gm);
}

/* ---------------------------------------------------------------- fs oprofile */

/* src/impl_sse/p7_fs_oprofile.c:222-296, minus striping.
 * esl_sse_expf for emissions and core transitions; libm expf for specials. */
BO_FS_OPROFILE *bo_fs_oprofile_convert(const BO_FS_PROFILE *gm)
{
  BO_FS_OPROFILE *om = calloc(1, sizeof(BO_FS_OPROFILE));
  int M = gm->M, c, k, z;
  static const int gmap[8] = { BO_P_BM, BO_P_MM, BO_P_IM, BO_P_DM, BO_P_MD, BO_P_MI, BO_P_II, BO_P_DD };

  if (!om) return NULL;
  om->M = M; om->L = gm->L; om->mode = gm->mode; om->nj = gm->nj;
  om->codon_lengths = gm->codon_lengths; om->maxcodons = gm->maxcodons;
  om->nrows = gm->maxcodons + BO_KP;
  for (z = 0; z < 8; z++) om->evparam[z] = gm->evparam[z];
  om->rfv = malloc(sizeof(float) * (size_t) om->nrows * (M + 1));
  om->tfv = malloc(sizeof(float) * (size_t) 8 * (M + 1));
  if (!om->rfv || !om->tfv) { bo_fs_oprofile_destroy(om); return NULL; }

  for (c = 0; c < om->nrows; c++) {
    om->rfv[(size_t) c * (M+1)] = 0.0f;
    for (k = 1; k <= M; k++)
      om->rfv[(size_t) c * (M+1) + k] = bo_cephes_expf(gm->rsc[(size_t) c * (M+1) + k]);
  }
  for (z = 0; z < 8; z++) {
    for (k = 0; k < M; k++)
      om->tfv[(size_t) z * (M+1) + k] = bo_cephes_expf(gm->tsc[k * BO_P_NTRANS + gmap[z]]);
    om->tfv[(size_t) z * (M+1) + M] = 0.0f;   /* (kb + z*nq < M) ? ... : -inf  -> 0 */
  }
  /* "straight" transitions never index node 0 in the striped layout */
  om->tfv[(size_t) BO_T_MD * (M+1)] = 0.0f;
  om->tfv[(size_t) BO_T_MI * (M+1)] = 0.0f;
  om->tfv[(size_t) BO_T_II * (M+1)] = 0.0f;
  om->tfv[(size_t) BO_T_DD * (M+1)] = 0.0f;

  om->xf[BO_X_E][BO_O_LOOP] = expf(gm->xsc[BO_X_E][BO_X_LOOP]);
  om->xf[BO_X_E][BO_O_MOVE] = expf(gm->xsc[BO_X_E][BO_X_MOVE]);
  om->xf[BO_X_N][BO_O_LOOP] = expf(gm->xsc[BO_X_N][BO_X_LOOP]);
  om->xf[BO_X_N][BO_O_MOVE] = expf(gm->xsc[BO_X_N][BO_X_MOVE]);
  om->xf[BO_X_C][BO_O_LOOP] = expf(gm->xsc[BO_X_C][BO_X_LOOP]);
  om->xf[BO_X_C][BO_O_MOVE] = expf(gm->xsc[BO_X_C][BO_X_MOVE]);
  om->xf[BO_X_J][BO_O_LOOP] = expf(gm->xsc[BO_X_J][BO_X_LOOP]);
  om->xf[BO_X_J][BO_O_MOVE] = expf(gm->xsc[BO_X_J][BO_X_MOVE]);
  return om;
}

void bo_fs_oprofile_destroy(BO_FS_OPROFILE *om)
{
  if (!om) return;
  free(om->rfv); free(om->tfv); free(om);
}

/* src/impl_sse/p7_fs_oprofile.c:636-651 */
void bo_fs_oprofile_ReconfigLength(BO_FS_OPROFILE *om, int L)
{
  float pmove = (2.0f + om->nj) / ((float) L + 2.0f + om->nj);
  float ploop = 1.0f - pmove;
  om->xf[BO_X_N][BO_O_LOOP] = om->xf[BO_X_C][BO_O_LOOP] = om->xf[BO_X_J][BO_O_LOOP] = ploop;
  om->xf[BO_X_N][BO_O_MOVE] = om->xf[BO_X_C][BO_O_MOVE] = om->xf[BO_X_J][BO_O_MOVE] = pmove;
  om->L = L;
}
/* :713-722 */
void bo_fs_oprofile_ReconfigMultihit(BO_FS_OPROFILE *om, int L)
{
  om->xf[BO_X_E][BO_O_MOVE] = 0.5;
  om->xf[BO_X_E][BO_O_LOOP] = 0.5;
  om->nj = 1.0f;
  bo_fs_oprofile_ReconfigLength(om, L);
}
/* :734-743 */
void bo_fs_oprofile_ReconfigUnihit(BO_FS_OPROFILE *om, int L)
{
  om->xf[BO_X_E][BO_O_MOVE] = 1.0f;
  om->xf[BO_X_E][BO_O_LOOP] = 0.0f;
  om->nj = 0.0f;
  bo_fs_oprofile_ReconfigLength(om, L);
}
