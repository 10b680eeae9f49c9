/* msv_avx2.c -- ORACLE side (test / bench infrastructure only; see bath_oracle.h): an AVX2 build of p7_MSVFilter with its SSV shortcut.
 *
 * The reference's production filters are SIMD (src/impl_sse/msvfilter.c, ssvfilter.c; 16 byte lanes, 32 in impl_avx); neither compiles
 * here (Easel is absent, INSTALL:6-8).  The scalar restatement in filters.c (bo_MSVFilter = msvfilter.c:74-208 over ssvfilter.c:831-925)
 * stays the CHECKER; this file runs the same byte arithmetic 32 nodes at a time so that the CPU arm of bench.py's search metric spends
 * its MSV time -- three quarters of the CPU search -- the way a SIMD build does.  Integer saturating arithmetic: the results are the
 * scalar ones bit for bit (tests/test_oracle_simd.py).
 *
 * Layout: nodes in natural order, not striped -- the k-1 look-back is an unaligned load one byte to the left. */
#define _POSIX_C_SOURCE 200112L
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <immintrin.h>
#include "bath_oracle.h"

#define PADL 32
#define V2   __attribute__((target("avx2")))

struct bo_msv_simd_s { int M, stride; uint8_t *rbp; int8_t *sbp; };

int bo_msv_simd_supported(void) { return __builtin_cpu_supports("avx2"); }

void bo_msv_simd_destroy(bo_msv_simd *im) { if (im) { free(im->rbp); free(im->sbp); free(im); } }

/* rbp[x][PADL + k] = rbv[x][k] (255 in the padding: a cell there is 0 in the MSV pass);
 * sbp[x][PADL + k] = ((127 + bias) -sat rbv[x][k]) ^ 127 read as signed (ssvfilter.c:750-757) */
bo_msv_simd *bo_msv_simd_create(const BO_OPROFILE *om)
{
  bo_msv_simd *im = calloc(1, sizeof *im);
  int M = om->M, x, k;
  if (!im) return NULL;
  im->M = M; im->stride = PADL + ((M + 31) / 32) * 32 + 32;
  im->rbp = malloc((size_t) BO_KP * im->stride);
  im->sbp = malloc((size_t) BO_KP * im->stride);
  if (!im->rbp || !im->sbp) { bo_msv_simd_destroy(im); return NULL; }
  memset(im->rbp, 255, (size_t) BO_KP * im->stride);
  memset(im->sbp, 0, (size_t) BO_KP * im->stride);
  for (x = 0; x < BO_KP; x++)
    for (k = 1; k <= M; k++) {
      uint8_t rb = om->rbv[(size_t) x * (M + 1) + k];
      int     d  = (int) (uint8_t) (om->bias_b + 127) - (int) rb;
      im->rbp[(size_t) x * im->stride + PADL + k] = rb;
      im->sbp[(size_t) x * im->stride + PADL + k] = (int8_t) ((uint8_t) (d < 0 ? 0 : d) ^ 127);
    }
  return im;
}

V2 static inline uint8_t hmax_epu8(__m256i v)
{
  __m128i m = _mm_max_epu8(_mm256_castsi256_si128(v), _mm256_extracti128_si256(v, 1));
  m = _mm_max_epu8(m, _mm_srli_si128(m, 8));
  m = _mm_max_epu8(m, _mm_srli_si128(m, 4));
  m = _mm_max_epu8(m, _mm_srli_si128(m, 2));
  m = _mm_max_epu8(m, _mm_srli_si128(m, 1));
  return (uint8_t) _mm_cvtsi128_si32(m);
}

/* byte mask of the last vector: 0xFF where k > M */
V2 static inline __m256i tail_mask(int M, int nv)
{
  uint8_t m[32];
  int j, k0 = 1 + 32 * (nv - 1);
  for (j = 0; j < 32; j++) m[j] = (k0 + j > M) ? 0xFF : 0;
  return _mm256_loadu_si256((const __m256i *) m);
}

/* get_xE (ssvfilter.c:831-874; filters.c ssv_get_xE): signed saturating diagonals from -128, unsigned running maximum */
V2 static uint8_t ssv_get_xE_simd(const bo_msv_simd *im, const uint8_t *dsq, int L, int8_t *prev, int8_t *cur)
{
  const int M = im->M, nv = (M + 31) / 32;
  const __m256i floorv = _mm256_set1_epi8((char) -128), tmask = tail_mask(M, nv);
  __m256i xEv = floorv;
  int i, v;
  memset(prev, 0x80, (size_t) im->stride); memset(cur, 0x80, (size_t) im->stride);
  for (i = 1; i <= L; i++) {
    const int8_t *sb = im->sbp + (size_t) dsq[i] * im->stride;
    int8_t *tmp;
    for (v = 0; v < nv; v++) {
      const int k0 = 1 + 32 * v;
      __m256i p = _mm256_loadu_si256((const __m256i *) (prev + PADL + k0 - 1));
      __m256i s = _mm256_loadu_si256((const __m256i *) (sb + PADL + k0));
      __m256i c = _mm256_subs_epi8(p, s);
      if (v == nv - 1) c = _mm256_blendv_epi8(c, floorv, tmask);
      xEv = _mm256_max_epu8(xEv, c);
      _mm256_storeu_si256((__m256i *) (cur + PADL + k0), c);
    }
    tmp = prev; prev = cur; cur = tmp;
  }
  return hmax_epu8(xEv);
}

static inline uint8_t u8_adds(uint8_t a, uint8_t b) { int s = (int) a + b; return (uint8_t)(s > 255 ? 255 : s); }
static inline uint8_t u8_subs(uint8_t a, uint8_t b) { int s = (int) a - b; return (uint8_t)(s < 0 ? 0 : s); }
static inline uint8_t u8_max(uint8_t a, uint8_t b)  { return a > b ? a : b; }

/* p7_MSVFilter (msvfilter.c:74-208) with the SSV shortcut first (:102-104; ssvfilter.c:876-925), as filters.c bo_MSVFilter */
V2 int bo_MSVFilter_simd(const bo_msv_simd *im, const uint8_t *dsq, int L, const BO_OPROFILE *om, float *ret_sc)
{
  const int M = im->M, nv = (M + 31) / 32;
  uint8_t buf0[PADL + 1024 + 64], buf1[PADL + 1024 + 64];
  uint8_t *prev = buf0, *cur = buf1, *tmp;
  uint8_t xJ, xB, xE;
  const uint8_t tjbm = (uint8_t)((int8_t) om->tjb_b + (int8_t) om->tbm_b);
  int i, v;
  if (M > 1024 || im->stride > (int) sizeof buf0) return BO_EINVAL;

  if (!(om->tjb_b + om->tbm_b + om->tec_b + om->bias_b >= 127)) {        /* bo_SSVFilter */
    uint16_t e = ssv_get_xE_simd(im, dsq, L, (int8_t *) buf0, (int8_t *) buf1), j;
    int answered = 1;
    if (e >= 255 - om->bias_b) {
      *ret_sc = INFINITY;
      if (om->base_b - om->tjb_b - om->tbm_b < 128) answered = 0; else return BO_ERANGE;
    } else {
      e += om->base_b - om->tjb_b - om->tbm_b;
      e -= 128;
      if (e >= 255 - om->bias_b) { *ret_sc = INFINITY; return BO_ERANGE; }
      j = e - om->tec_b;
      if (j > om->base_b) answered = 0;
      else {
        *ret_sc  = ((float) (j - om->tjb_b) - (float) om->base_b);
        *ret_sc /= om->scale_b;
        *ret_sc -= 3.0;
        return BO_OK;
      }
    }
    (void) answered;
  }

  memset(buf0, 0, (size_t) im->stride); memset(buf1, 0, (size_t) im->stride);
  xJ = 0;
  xB = u8_subs(om->base_b, tjbm);
  {
    const __m256i biasv = _mm256_set1_epi8((char) om->bias_b);
    for (i = 1; i <= L; i++) {
      const uint8_t *rb = im->rbp + (size_t) dsq[i] * im->stride;
      const __m256i xBv = _mm256_set1_epi8((char) xB);
      __m256i xEv = _mm256_setzero_si256();
      for (v = 0; v < nv; v++) {
        const int k0 = 1 + 32 * v;
        __m256i sv = _mm256_max_epu8(_mm256_loadu_si256((const __m256i *) (prev + PADL + k0 - 1)), xBv);
        sv = _mm256_adds_epu8(sv, biasv);
        sv = _mm256_subs_epu8(sv, _mm256_loadu_si256((const __m256i *) (rb + PADL + k0)));
        xEv = _mm256_max_epu8(xEv, sv);
        _mm256_storeu_si256((__m256i *) (cur + PADL + k0), sv);
      }
      xE = hmax_epu8(xEv);
      if (u8_adds(xE, om->bias_b) == 255) { *ret_sc = INFINITY; return BO_ERANGE; }
      xE = u8_subs(xE, om->tec_b);
      xJ = u8_max(xJ, xE);
      xB = u8_max(om->base_b, xJ);
      xB = u8_subs(xB, tjbm);
      tmp = prev; prev = cur; cur = tmp;
    }
  }
  *ret_sc  = ((float) (xJ - om->tjb_b) - (float) om->base_b);
  *ret_sc /= om->scale_b;
  *ret_sc -= 3.0;
  return BO_OK;
}
