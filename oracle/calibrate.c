/* calibrate.c -- ORACLE (test infrastructure only; never linked into the product).
 *
 * E-value calibration of a query model by brief simulation, the part of bathbuild / bathconvert / bathfetch that runs the
 * hot-path kernels on random sequences:
 *   p7_Calibrate                 src/evalues.c:64-183   (order of the five simulations, one generator threaded through all)
 *   p7_Lambda                    src/evalues.c:243-250  + p7_MeanMatchRelativeEntropy, src/modelstats.c:79-97
 *   p7_MSVMu / p7_ViterbiMu      src/evalues.c:297-340, :366-411
 *   p7_Tau                       src/evalues.c:536-581
 *   p7_fs_Tau_3codons / _5codons src/evalues.c:607-680, :703-776
 *   p7_codontable_Create/GetCodon src/hmmer.c:197-243, :257-273
 *
 * PARITY PINNED: with the builder's default seed (42) this reproduces the five STATS lines of the shipped models
 * (tutorial/AMP_N.bhmm:15-19, PTH2.bhmm:20-24, tRNA-synthetases.bhmm x3, ...) to the four decimals they are printed with
 * (tests/test_calibration.py), which pins in one go Easel's fast generator and esl_rnd_FChoose as restated in
 * fs_stotrace.c, the MSV and Viterbi filters, the protein Forward parser and both frameshift Forward recursions against
 * numbers the reference itself produced.
 *
 * Easel pieces restated from its published sources (Easel is absent from /root/reference):
 *   esl_rsq_xfIID (one esl_rnd_FChoose per residue), esl_rnd_Roll (esl_random() * n), esl_vec_FRelEntropy,
 *   esl_gumbel_FitComplete (Newton-Raphson on Lawless' eq. 4.1.6, tolerance 1e-5, then eq. 4.1.5),
 *   esl_gumbel_FitCompleteLoc (eq. 4.1.5 at a known lambda), esl_gumbel_invcdf. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "bath_oracle.h"

#define LOG2 0.69314718055994529

/* esl_rnd_FChoose over the background frequencies */
static int choose_residue(BO_RNG *r, const float *p, int n)
{
  double roll = bo_random(r), sum = 0.0;
  int    i;
  for (i = 0; i < n; i++) { sum += p[i]; if (roll < sum) return i; }
  do { i = (int)(bo_random(r) * n); } while (p[i] == 0.0f);
  return i;
}

static void random_protein(BO_RNG *r, const float *f, int L, uint8_t *dsq)
{
  int i;
  dsq[0] = dsq[L + 1] = 255;
  for (i = 1; i <= L; i++) dsq[i] = (uint8_t) choose_residue(r, f, BO_K);
}

/* p7_codontable_Create: codons of each amino acid in the order x, y, z run (src/hmmer.c:222-234) */
typedef struct { uint8_t nt[BO_K][6][3]; int n[BO_K]; } codon_table;
static int codon_table_fill(codon_table *t, int ct)
{
  const uint8_t *basic = bo_gencode_basic(ct);
  int c;
  if (!basic) return BO_EINVAL;
  memset(t, 0, sizeof *t);
  for (c = 0; c < 64; c++) {
    int a = basic[c];
    if (a < BO_K && t->n[a] < 6) {
      t->nt[a][t->n[a]][0] = (uint8_t)(c >> 4); t->nt[a][t->n[a]][1] = (uint8_t)((c >> 2) & 3); t->nt[a][t->n[a]][2] = (uint8_t)(c & 3);
      t->n[a]++;
    }
  }
  return BO_OK;
}

/* the reverse translation loop of p7_fs_Tau_*codons (src/evalues.c:638-647) */
static void random_coding_dna(BO_RNG *r, const float *f, const codon_table *t, int L, uint8_t *amino, uint8_t *dna)
{
  int a, j = 1;
  random_protein(r, f, L, amino);
  dna[0] = dna[3 * L + 1] = 255;
  for (a = 1; a <= L; a++, j += 3) {
    int x = (int)(bo_random(r) * t->n[amino[a]]);        /* esl_rnd_Roll */
    memcpy(dna + j, t->nt[amino[a]][x], 3);
  }
}

double bo_gumbel_FitCompleteLoc(const double *x, int n, double lambda)
{
  double esum = 0.0;
  int    i;
  for (i = 0; i < n; i++) esum += exp(-lambda * x[i]);
  return -log(esum / n) / lambda;
}

void bo_gumbel_FitComplete(const double *x, int n, double *ret_mu, double *ret_lambda)
{
  double mean = 0.0, var = 0.0, lambda, fx, dfx;
  int    i, it;
  for (i = 0; i < n; i++) mean += x[i];
  mean /= n;
  for (i = 0; i < n; i++) var += (x[i] - mean) * (x[i] - mean);
  var /= (n - 1);
  lambda = 3.14159265358979323846 / sqrt(6.0 * var);
  for (it = 0; it < 100; it++) {
    double esum = 0., xesum = 0., xxesum = 0., xsum = 0.;
    for (i = 0; i < n; i++) {
      double e = exp(-lambda * x[i]);
      xsum += x[i]; esum += e; xesum += x[i] * e; xxesum += x[i] * x[i] * e;
    }
    fx  = (1.0 / lambda) - (xsum / n) + (xesum / esum);
    dfx = ((xesum / esum) * (xesum / esum)) - (xxesum / esum) - (1.0 / (lambda * lambda));
    if (fabs(fx) < 1e-5) break;
    lambda -= fx / dfx;
    if (lambda <= 0.0) lambda = 0.001;
  }
  *ret_lambda = lambda;
  *ret_mu = bo_gumbel_FitCompleteLoc(x, n, lambda);
}

/* src/evalues.c:564-568 */
static double tau_of(const double *xv, int n, double lambda, double tailp)
{
  double gmu, glam;
  bo_gumbel_FitComplete(xv, n, &gmu, &glam);
  return (gmu - log(-log(1.0 - tailp)) / glam) + log(tailp) / lambda;
}

/* p7_Lambda (src/evalues.c:243-250): log 2 + 1.44 / (M H), H = mean relative entropy of the match emissions in bits */
double bo_Lambda(const BO_HMM *hmm, const BO_BG *bg)
{
  double KL = 0.0;
  int    k, x;
  for (k = 1; k <= hmm->M; k++) {
    float kl = 0.0f;
    const float *p = hmm->mat + (size_t) k * BO_K;
    for (x = 0; x < BO_K; x++) if (p[x] > 0.0f) kl += p[x] * log(p[x] / bg->f[x]);
    KL += kl / LOG2;
  }
  KL /= (double) hmm->M;
  return LOG2 + 1.44 / ((double) hmm->M * KL);
}

/* p7_Calibrate with the frameshift branch on (cfg_b->fs): out[8] in evparam order.  lambda <= 0: p7_Lambda of the model.
 * which_mask selects the simulations to run (bit 0 MSV, 1 Viterbi, 2 Forward, 3 FS3, 4 FS5); the generator is advanced past
 * skipped simulations so that every value is the one the full sequence of calls yields.
 * convert_flow = 1: what bathconvert / bathfetch do to a model that has protein statistics but no frameshift ones
 * (src/bathconvert.c:128,157-161; src/bathfetch.c:295,321-325): only the two frameshift simulations, on a generator that
 * is created once per run (seed 42) and keeps running from one model of the file to the next -- *rng_x carries that state
 * (0 on entry = fresh generator) and receives it back. */
int bo_Calibrate(const BO_HMM *hmm, BO_BG *bg, BO_OPROFILE *om, BO_FS_OPROFILE *om_fs3, BO_FS_OPROFILE *om_fs5, int ct,
                 uint32_t seed, double lambda, int which_mask, int convert_flow, uint32_t *rng_x, double out[8])
{
  const int EmL = 200, EmN = 200, EvL = 200, EvN = 200, EfL = 100, EfN = 200;
  const double Eft = 0.04;
  BO_RNG   rng;
  codon_table tbl;
  uint8_t *dsq = malloc(EmL + EvL + 2), *dna = malloc(3 * EfL + 2);
  double  *xv = malloc(sizeof(double) * 256);
  float    sc, nullsc;
  int      i, st, status = BO_OK;

  if (!dsq || !dna || !xv) { status = BO_EMEM; goto done; }
  if ((status = codon_table_fill(&tbl, ct)) != BO_OK) goto done;
  if (lambda <= 0.0) lambda = bo_Lambda(hmm, bg);
  for (i = 0; i < 8; i++) out[i] = -99999.0;
  out[1] = out[3] = out[5] = lambda;
  bo_rng_init(&rng, seed);
  if (convert_flow && rng_x && *rng_x) rng.x = *rng_x;
  if (convert_flow) goto frameshift;

  /* p7_MSVMu: overflow counts as the largest representable score */
  bo_oprofile_ReconfigLength(om, EmL); bo_bg_SetLength(bg, EmL);
  for (i = 0; i < EmN; i++) {
    random_protein(&rng, bg->f, EmL, dsq);
    if (!(which_mask & 1)) continue;
    nullsc = bo_bg_NullOne(bg, EmL);
    st = bo_MSVFilter(dsq, EmL, om, &sc);
    if (st == BO_ERANGE) sc = (255 - om->base_b) / om->scale_b; else if (st != BO_OK) { status = st; goto done; }
    xv[i] = (sc - nullsc) / LOG2;
  }
  if (which_mask & 1) out[0] = bo_gumbel_FitCompleteLoc(xv, EmN, lambda);

  /* p7_ViterbiMu */
  bo_oprofile_ReconfigLength(om, EvL); bo_bg_SetLength(bg, EvL);
  for (i = 0; i < EvN; i++) {
    random_protein(&rng, bg->f, EvL, dsq);
    if (!(which_mask & 2)) continue;
    nullsc = bo_bg_NullOne(bg, EvL);
    st = bo_ViterbiFilter(dsq, EvL, om, &sc);
    if (st == BO_ERANGE) sc = (32767.0 - om->base_w) / om->scale_w; else if (st != BO_OK) { status = st; goto done; }
    xv[i] = (sc - nullsc) / LOG2;
  }
  if (which_mask & 2) out[2] = bo_gumbel_FitCompleteLoc(xv, EvN, lambda);

  /* p7_Tau */
  bo_oprofile_ReconfigLength(om, EfL); bo_bg_SetLength(bg, EfL);
  for (i = 0; i < EfN; i++) {
    random_protein(&rng, bg->f, EfL, dsq);
    if (!(which_mask & 4)) continue;
    if ((st = bo_ForwardParser(dsq, EfL, om, &sc)) != BO_OK) { status = st; goto done; }
    xv[i] = (sc - bo_bg_NullOne(bg, EfL)) / LOG2;
  }
  if (which_mask & 4) out[4] = tau_of(xv, EfN, lambda, Eft);

frameshift:
  /* p7_fs_Tau_3codons: the length model is set with the AMINO length (src/evalues.c:628), a sequence whose score
   * overflows is drawn again (:649) */
  bo_fs_oprofile_ReconfigLength(om_fs3, EfL); bo_bg_SetLength(bg, EfL);
  {
    BO_MX *ox = bo_mx_create(om_fs3->M, 3 * EfL, 0);
    if (!ox) { status = BO_EMEM; goto done; }
    for (i = 0; i < EfN; i++) {
      random_coding_dna(&rng, bg->f, &tbl, EfL, dsq, dna);
      if (!(which_mask & 8)) continue;
      st = bo_ForwardParser_Frameshift_3Codons(dna, 3 * EfL, om_fs3, ox, &sc);
      if (st == BO_ERANGE) { i--; continue; }
      if (st != BO_OK) { bo_mx_destroy(ox); status = st; goto done; }
      xv[i] = (sc - bo_bg_fs_NullOne(bg, EfL)) / LOG2;
    }
    bo_mx_destroy(ox);
  }
  if (which_mask & 8) out[6] = tau_of(xv, EfN, lambda, Eft);

  /* p7_fs_Tau_5codons (the parser's score equals the full-matrix Forward's) */
  bo_fs_oprofile_ReconfigLength(om_fs5, EfL); bo_bg_SetLength(bg, EfL);
  {
    BO_MX *ox = bo_mx_create(om_fs5->M, 3 * EfL, BO_NSCELLS_FS);
    if (!ox) { status = BO_EMEM; goto done; }
    for (i = 0; i < EfN; i++) {
      random_coding_dna(&rng, bg->f, &tbl, EfL, dsq, dna);
      if (!(which_mask & 16)) continue;
      st = bo_Forward_Frameshift(dna, 3 * EfL, om_fs5, ox, &sc);
      if (st == BO_ERANGE) { i--; continue; }
      if (st != BO_OK) { bo_mx_destroy(ox); status = st; goto done; }
      xv[i] = (sc - bo_bg_fs_NullOne(bg, EfL)) / LOG2;
    }
    bo_mx_destroy(ox);
  }
  if (which_mask & 16) out[7] = tau_of(xv, EfN, lambda, Eft);
  if (rng_x) *rng_x = rng.x;

done:
  free(dsq); free(dna); free(xv);
  return status;
}
