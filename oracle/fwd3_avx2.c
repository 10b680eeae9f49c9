/* fwd3_avx2.c -- ORACLE side (test / bench infrastructure only; see bath_oracle.h): an AVX2 + FMA build of the 3-codon Forward parser.
 *
 * The reference's production build of p7_ForwardParser_Frameshift_3Codons is SIMD (src/impl_sse/fwdback_fs.c:97-533, the AVX2 twin in
 * src/impl_avx); neither compiles here (Easel is absent, INSTALL:6-8).  The scalar restatement in fs_fwdback.c stays the CHECKER; this
 * file is the same recurrence, row for row (fs_fwdback.c:72-137 = fwdback_fs.c:340-505), vectorised eight nodes at a time so that
 * bench.py's CPU arm is a SIMD program on all host cores and not a scalar one.  It is parity-tested against the scalar oracle
 * (tests/test_oracle_simd.py); scores differ in the last bits because the products are fused here.
 *
 * Layout: nodes in natural order, not striped -- with unaligned loads the k-1 look-back is a load at offset -1, and the one serial
 * part of a row, the delete chain D(k) = D(k-1) tDD(k-1) + M(k-1) tMD(k-1), is a first-order linear recurrence that is scanned
 * inside each vector with constant multipliers (three permute + FMA steps) and chained from vector to vector through one carry.
 * Scores only (no X rows kept): what the window filter stage needs (src/p7_pipeline.c:1450).
 */
#define _POSIX_C_SOURCE 200112L
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <immintrin.h>
#include "bath_oracle.h"

#define PADL 8                      /* zero floats in front of node 1 (the k-1 loads of the first vector) */
#define V8   __attribute__((target("avx2,fma")))

int bo_fwd3_simd_supported(void) { return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma"); }

typedef struct {
  int    M, nv, stride;             /* nv vectors of 8 nodes; stride floats per row = PADL + 8 nv + 8 */
  float *rf;                        /* [338][stride]  emission odds, node k at [PADL + k - 1]... see AT() */
  float *tbm1, *tmm1, *tim1, *tdm1; /* transition odds out of node k-1, stored at node k */
  float *tmd1;                      /* tMD(k-1) at node k */
  float *tmi, *tii;                 /* at node k */
  float *a1, *a2, *a4, *cp;         /* delete-chain scan multipliers per vector lane */
  float  xf[4][2];
} simd_image;

#define AT(row, k) ((row) + PADL + (k) - 1)        /* node k (1-based) of a padded row */

static float *alloc32(size_t nfloats) { void *p = NULL; return posix_memalign(&p, 32, sizeof(float) * nfloats) == 0 ? (float *) p : NULL; }
static float *zrow(int stride) { float *p = alloc32((size_t) stride); memset(p, 0, sizeof(float) * (size_t) stride); return p; }

static simd_image *image_create(const BO_FS_OPROFILE *om)
{
  simd_image *im = calloc(1, sizeof(simd_image));
  int M = om->M, k, c, v, l;
  im->M = M; im->nv = (M + 7) / 8; im->stride = PADL + 8 * im->nv + 8;
  im->rf = alloc32((size_t) BO_MAXCODONS3 * im->stride);
  memset(im->rf, 0, sizeof(float) * (size_t) BO_MAXCODONS3 * im->stride);
  for (c = 0; c < BO_MAXCODONS3; c++)
    for (k = 1; k <= M; k++) *AT(im->rf + (size_t) c * im->stride, k) = om->rfv[(size_t) c * (M + 1) + k];
  im->tbm1 = zrow(im->stride); im->tmm1 = zrow(im->stride); im->tim1 = zrow(im->stride); im->tdm1 = zrow(im->stride);
  im->tmd1 = zrow(im->stride); im->tmi = zrow(im->stride); im->tii = zrow(im->stride);
  im->a1 = zrow(im->stride); im->a2 = zrow(im->stride); im->a4 = zrow(im->stride); im->cp = zrow(im->stride);
#define T(t, k) (om->tfv[(size_t)(t) * (M + 1) + (k)])
  for (k = 1; k <= M; k++) {
    *AT(im->tbm1, k) = T(BO_T_BM, k - 1); *AT(im->tmm1, k) = T(BO_T_MM, k - 1);
    *AT(im->tim1, k) = T(BO_T_IM, k - 1); *AT(im->tdm1, k) = T(BO_T_DM, k - 1);
    *AT(im->tmd1, k) = (k >= 2) ? T(BO_T_MD, k - 1) : 0.0f;
    *AT(im->tmi, k)  = T(BO_T_MI, k);     *AT(im->tii, k)  = T(BO_T_II, k);
  }
  /* D(k) = a(k) D(k-1) + b(k), a(k) = tDD(k-1) (a(1) = 0: node 1 has no delete state) */
  for (v = 0; v < im->nv; v++)
    for (l = 0; l < 8; l++) {
      int   kk = 8 * v + l + 1, z;
      float a[8], p;
      for (z = 0; z < 8; z++) { int kz = kk - z; a[z] = (kz >= 2 && kz <= M) ? T(BO_T_DD, kz - 1) : 0.0f; }   /* a(kk), a(kk-1), ... */
      *AT(im->a1, kk) = (l >= 1) ? a[0] : 0.0f;
      *AT(im->a2, kk) = (l >= 2) ? a[0] * a[1] : 0.0f;
      *AT(im->a4, kk) = (l >= 4) ? a[0] * a[1] * a[2] * a[3] : 0.0f;
      for (p = 1.0f, z = 0; z <= l; z++) p *= a[z];
      *AT(im->cp, kk) = p;
    }
#undef T
  memcpy(im->xf, om->xf, sizeof im->xf);
  return im;
}

static void image_destroy(simd_image *im)
{
  if (!im) return;
  free(im->rf); free(im->tbm1); free(im->tmm1); free(im->tim1); free(im->tdm1); free(im->tmd1); free(im->tmi); free(im->tii);
  free(im->a1); free(im->a2); free(im->a4); free(im->cp); free(im);
}

typedef struct { float *mm[4], *im_[4], *dm[4], *iv[3], *mem; } simd_rows;

static int rows_create(simd_rows *R, int stride)
{
  int r;
  R->mem = alloc32((size_t) stride * 15);
  if (!R->mem) return BO_EMEM;
  for (r = 0; r < 4; r++) { R->mm[r] = R->mem + (size_t)(3 * r) * stride; R->im_[r] = R->mem + (size_t)(3 * r + 1) * stride; R->dm[r] = R->mem + (size_t)(3 * r + 2) * stride; }
  for (r = 0; r < 3; r++) R->iv[r] = R->mem + (size_t)(12 + r) * stride;
  return BO_OK;
}

static inline int nuc3(uint8_t d) { return (d < BO_MAXNUC) ? d : BO_MAXCODONS3; }
static inline int pmod(int a, int n) { return ((a % n) + n) % n; }

V8 static inline float hsum8(__m256 v)
{
  __m128 s = _mm_add_ps(_mm256_castps256_ps128(v), _mm256_extractf128_ps(v, 1));
  s = _mm_add_ps(s, _mm_movehl_ps(s, s));
  s = _mm_add_ss(s, _mm_shuffle_ps(s, s, 1));
  return _mm_cvtss_f32(s);
}

/* one window; pmove / ploop: the N/J/C odds of the window's length model (p7_fs_oprofile_ReconfigLength) */
V8 static int fwd3_window(const simd_image *im, simd_rows *R, const uint8_t *dsq, int L, float pmove, float ploop, float *opt_sc)
{
  const int   nv = im->nv, stride = im->stride;
  const float tEL = im->xf[BO_X_E][BO_O_LOOP], tEM = im->xf[BO_X_E][BO_O_MOVE];
  float  xN, xE, xB, xC, xJ, xNb[4], xBb[4], xJb[4], xCb[4];
  double totscale = 0.0;
  int    i, r, v, u, vv, w, x;
  const __m256i sh1 = _mm256_setr_epi32(0, 0, 1, 2, 3, 4, 5, 6), sh2 = _mm256_setr_epi32(0, 0, 0, 1, 2, 3, 4, 5),
                sh4 = _mm256_setr_epi32(0, 0, 0, 0, 0, 1, 2, 3), last = _mm256_set1_epi32(7);

  if (L < 3) return BO_EINVAL;
  memset(R->mem, 0, sizeof(float) * (size_t) stride * 15);
  for (r = 0; r < 4; r++) xNb[r] = xBb[r] = xJb[r] = xCb[r] = 0.0f;
  xNb[0] = xNb[1] = 1.0f;
  xBb[0] = xBb[1] = pmove;
  u = vv = BO_MAXCODONS3;
  w = nuc3(dsq[1]);
  x = nuc3(dsq[2]);

  for (i = 2; i <= L; i++) {
    const int curr = i % 4, prev2 = pmod(i - 2, 4), prev3 = pmod(i - 3, 4);
    float *mmc = R->mm[curr], *imc = R->im_[curr], *dmc = R->dm[curr];
    const float *mm2 = R->mm[prev2], *im2 = R->im_[prev2], *dm2 = R->dm[prev2], *mm3 = R->mm[prev3], *im3 = R->im_[prev3];
    float *iv2 = R->iv[i % 3];
    const float *iv3 = R->iv[pmod(i - 1, 3)], *iv4 = R->iv[pmod(i - 2, 3)];
    const float *r2, *r3, *r4;
    int c2, c3, c4;
    __m256 xEv = _mm256_setzero_ps(), xDv = _mm256_setzero_ps(), carry = _mm256_setzero_ps(), mlast = _mm256_setzero_ps();
    const __m256 xB2 = _mm256_set1_ps(xBb[prev2]);

    if (i > 2) { u = vv; vv = w; w = x; x = nuc3(dsq[i]); }
    c2 = BO_CODON2_FS3(w, x);         c2 = BO_MINIDX(c2, BO_DEGEN3_QC1);
    c3 = BO_CODON3_FS3(vv, w, x);     c3 = BO_MINIDX(c3, BO_DEGEN3_C);
    c4 = BO_CODON4_FS3(u, vv, w, x);  c4 = BO_MINIDX(c4, BO_DEGEN3_QC1);
    r2 = im->rf + (size_t) c2 * stride; r3 = im->rf + (size_t) c3 * stride; r4 = im->rf + (size_t) c4 * stride;

    for (v = 0; v < nv; v++) {
      const int k = 8 * v + 1, o = PADL + k - 1;
      __m256 sv, msv, b, s;
      sv  = _mm256_mul_ps(xB2, _mm256_loadu_ps(im->tbm1 + o));
      sv  = _mm256_fmadd_ps(_mm256_loadu_ps(mm2 + o - 1), _mm256_loadu_ps(im->tmm1 + o), sv);
      sv  = _mm256_fmadd_ps(_mm256_loadu_ps(im2 + o - 1), _mm256_loadu_ps(im->tim1 + o), sv);
      sv  = _mm256_fmadd_ps(_mm256_loadu_ps(dm2 + o - 1), _mm256_loadu_ps(im->tdm1 + o), sv);
      _mm256_storeu_ps(iv2 + o, sv);
      msv = _mm256_mul_ps(sv, _mm256_loadu_ps(r2 + o));
      msv = _mm256_fmadd_ps(_mm256_loadu_ps(iv3 + o), _mm256_loadu_ps(r3 + o), msv);     /* zero rows at i = 2 (:217-218) */
      msv = _mm256_fmadd_ps(_mm256_loadu_ps(iv4 + o), _mm256_loadu_ps(r4 + o), msv);
      xEv = _mm256_add_ps(xEv, msv);
      _mm256_storeu_ps(mmc + o, msv);
      _mm256_storeu_ps(imc + o, _mm256_fmadd_ps(_mm256_loadu_ps(im3 + o), _mm256_loadu_ps(im->tii + o),
                                                _mm256_mul_ps(_mm256_loadu_ps(mm3 + o), _mm256_loadu_ps(im->tmi + o))));
      /* delete chain of this vector: b(k) = M(i,k-1) tMD(k-1); scan inside the vector, then the carry of the vectors before */
      b = _mm256_blend_ps(_mm256_permutevar8x32_ps(msv, sh1), mlast, 1);          /* M(i,k-1): from registers, not from the stores just made */
      mlast = _mm256_permutevar8x32_ps(msv, last);
      b = _mm256_mul_ps(b, _mm256_loadu_ps(im->tmd1 + o));
      s = _mm256_fmadd_ps(_mm256_loadu_ps(im->a1 + o), _mm256_permutevar8x32_ps(b, sh1), b);
      s = _mm256_fmadd_ps(_mm256_loadu_ps(im->a2 + o), _mm256_permutevar8x32_ps(s, sh2), s);
      s = _mm256_fmadd_ps(_mm256_loadu_ps(im->a4 + o), _mm256_permutevar8x32_ps(s, sh4), s);
      s = _mm256_fmadd_ps(_mm256_loadu_ps(im->cp + o), carry, s);
      carry = _mm256_permutevar8x32_ps(s, last);
      _mm256_storeu_ps(dmc + o, s);
      xDv = _mm256_add_ps(xDv, s);
    }
    xE = hsum8(_mm256_add_ps(xEv, xDv));

    if (i == 2) { xN = 1.0f; xJ = xE * tEL; xC = xE * tEM; }
    else {
      xN = xNb[prev3] * ploop;
      xJ = xJb[prev3] * ploop + xE * tEL;
      xC = xCb[prev3] * ploop + xE * tEM;
    }
    xB = xN * pmove + xJ * pmove;

    if (xE > 1.0e4f) {
      const float sf = 1.0f / xE;
      const __m256 sfv = _mm256_set1_ps(sf);
      int z;
      xN *= sf; xJ *= sf; xC *= sf; xB *= sf;
      for (z = 0; z < stride * 15; z += 8) _mm256_store_ps(R->mem + z, _mm256_mul_ps(_mm256_load_ps(R->mem + z), sfv));
      for (r = 0; r < 4; r++) { xNb[r] *= sf; xBb[r] *= sf; xJb[r] *= sf; xCb[r] *= sf; }
      totscale += log(xE);
    }
    xNb[curr] = xN; xBb[curr] = xB; xJb[curr] = xJ; xCb[curr] = xC;
  }
  {
    const float xCtot = xCb[L % 4] + xCb[pmod(L - 1, 4)] * ploop + xCb[pmod(L - 2, 4)] * ploop;
    if (isnan(xCtot) || isinf(xCtot)) return BO_ERANGE;
    if (L > 2 && xCtot == 0.0f) { if (opt_sc) *opt_sc = -INFINITY; return BO_ERANGE; }
    if (opt_sc) *opt_sc = (float) totscale + logf(xCtot * pmove);
  }
  return BO_OK;
}

/* ---- batch over windows: one worker per thread, windows dealt by an atomic counter ---- */
typedef struct {
  const uint8_t *dsq; const int64_t *start; const int32_t *L; int n, maxL;
  const simd_image *im; float nj; float *sc; int32_t *status; int next;
} simd_job;

static void *simd_worker(void *arg)
{
  simd_job *job = arg;
  simd_rows R;
  uint8_t *sub = malloc((size_t) job->maxL + 2);
  _mm_setcsr(_mm_getcsr() | 0x8040);               /* flush-to-zero + denormals-are-zero, per thread, as impl_Init does (src/impl_sse/impl_sse.h:559-577; src/bathsearch.c:1235) */
  if (rows_create(&R, job->im->stride) != BO_OK) { free(sub); return NULL; }
  for (;;) {
    const int w = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);
    int L;
    float pmove, ploop;
    if (w >= job->n) break;
    L = job->L[w];
    sub[0] = BO_DSQ_SENTINEL;                       /* each window is handed over as its own sub-sequence (src/p7_pipeline.c:1376-1380) */
    memcpy(sub + 1, job->dsq + job->start[w], (size_t) L);
    sub[L + 1] = BO_DSQ_SENTINEL;
    pmove = (2.0f + job->nj) / ((float) (L / 3) + 2.0f + job->nj);     /* p7_fs_oprofile_ReconfigLength(om_fs3, L/3) (src/p7_pipeline.c:1449) */
    ploop = 1.0f - pmove;
    job->status[w] = fwd3_window(job->im, &R, sub, L, pmove, ploop, &job->sc[w]);
  }
  free(R.mem); free(sub);
  return NULL;
}

int bo_batch_ForwardParser_3Codons_simd(const uint8_t *dsq, const int64_t *start, const int32_t *L, int n,
                                        const BO_FS_OPROFILE *om, int nthreads, float *sc, int32_t *status)
{
  simd_job job;
  pthread_t *th;
  int t, maxL = 0;
  if (n < 1 || nthreads < 1 || om->codon_lengths != 3) return BO_EINVAL;
  if (!bo_fwd3_simd_supported()) return BO_EINVAL;
  for (t = 0; t < n; t++) if (L[t] > maxL) maxL = L[t];
  job.dsq = dsq; job.start = start; job.L = L; job.n = n; job.maxL = maxL; job.nj = om->nj; job.sc = sc; job.status = status; job.next = 0;
  job.im = image_create(om);
  th = malloc(sizeof(pthread_t) * (size_t) nthreads);
  for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, simd_worker, &job);
  for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  image_destroy((simd_image *) job.im);
  return BO_OK;
}
