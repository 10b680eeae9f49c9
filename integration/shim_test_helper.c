/* shim_test_helper.c -- TEST INFRASTRUCTURE for impl_cuda_shim.c: builds the reference's striped structures from plain arrays.
 *
 * shimtest_make_oprofile stripes plain tables the way fs_fb_conversion does (src/impl_sse/p7_fs_oprofile.c:222-296): emission row c,
 * stripe q, lane z holds node k = q + 1 + z Q; transition vectors are interleaved per stripe as {BM,MM,IM,DM,MD,MI,II}, the first
 * four taken at the source node k-1 and valid while k-1+zQ < M, the last three at node k and valid while k+zQ < M; Q vectors of DD
 * follow (src/impl_sse/p7_fs_oprofile.c:254-284).  Lanes past the model hold 0 (expf(-inf)).
 * shimtest_make_omx allocates a P7_OMX with the X rows the parsers write (src/impl_sse/p7_omx.c: xmx[(allocXR) * p7X_NXCELLS]). */
#include <stdlib.h>
#include <string.h>
#include "bath_impl_standin.h"

static ESL_ALPHABET shimtest_amino = { 3, 20, 29 };

P7_FS_OPROFILE *shimtest_make_oprofile(int which, int M, int nrows, const float *rfv, const float *tfv, const float xf[8])
{
  P7_FS_OPROFILE *om = calloc(1, sizeof(P7_FS_OPROFILE));
  const int Q = p7O_NQF(M);
  union { __m128 v; float x[4]; } tmp;
  int c, q, z, t, j = 0;
  om->M = M; om->allocM = M; om->allocQ4 = Q; om->codon_lengths = which; om->abc = &shimtest_amino; om->L = 100; om->nj = 1.0f; om->mode = 1;
  om->rfv_mem = aligned_alloc(16, sizeof(__m128) * (size_t) nrows * Q);
  om->tfv_mem = aligned_alloc(16, sizeof(__m128) * (size_t) p7O_NTRANS * Q);
  om->rfv = malloc(sizeof(__m128 *) * (size_t) nrows);
  om->tfv = om->tfv_mem;
  for (c = 0; c < nrows; c++) {
    om->rfv[c] = om->rfv_mem + (size_t) c * Q;
    for (q = 0; q < Q; q++) {
      for (z = 0; z < 4; z++) { const int k = q + 1 + z * Q; tmp.x[z] = (k <= M) ? rfv[(size_t) c * (M + 1) + k] : 0.0f; }
      om->rfv[c][q] = tmp.v;
    }
  }
  for (q = 0; q < Q; q++)
    for (t = p7O_BM; t <= p7O_II; t++) {
      const int kb = (t <= p7O_DM) ? q : q + 1;
      for (z = 0; z < 4; z++) tmp.x[z] = (kb + z * Q < M) ? tfv[(size_t) t * (M + 1) + kb + z * Q] : 0.0f;
      om->tfv[j++] = tmp.v;
    }
  for (q = 0; q < Q; q++) {
    for (z = 0; z < 4; z++) tmp.x[z] = (q + 1 + z * Q < M) ? tfv[(size_t) p7O_DD * (M + 1) + q + 1 + z * Q] : 0.0f;
    om->tfv[j++] = tmp.v;
  }
  memcpy(om->xf, xf, sizeof(float) * 8);
  return om;
}

void shimtest_free_oprofile(P7_FS_OPROFILE *om)
{
  if (!om) return;
  free(om->rfv_mem); free(om->tfv_mem); free(om->rfv); free(om);
}

P7_OMX *shimtest_make_omx(int L)
{
  P7_OMX *ox = calloc(1, sizeof(P7_OMX));
  ox->allocXR = L + 1;
  ox->x_mem = calloc((size_t) (L + 1) * p7X_NXCELLS + 4, sizeof(float));
  ox->xmx = ox->x_mem;
  return ox;
}
float *shimtest_omx_xmx(P7_OMX *ox)      { return ox->xmx; }
float  shimtest_omx_totscale(P7_OMX *ox) { return ox->totscale; }
int    shimtest_omx_field(P7_OMX *ox, int which) { return which == 0 ? ox->M : which == 1 ? ox->L : ox->has_own_scales; }
void   shimtest_free_omx(P7_OMX *ox) { if (ox) { free(ox->x_mem); free(ox); } }
void   shimtest_set_length(P7_FS_OPROFILE *om, float pmove, float ploop)      /* p7_fs_oprofile_ReconfigLength's effect on xf (p7_fs_oprofile.c:636-651) */
{
  om->xf[p7O_N][p7O_MOVE] = om->xf[p7O_C][p7O_MOVE] = om->xf[p7O_J][p7O_MOVE] = pmove;
  om->xf[p7O_N][p7O_LOOP] = om->xf[p7O_C][p7O_LOOP] = om->xf[p7O_J][p7O_LOOP] = ploop;
}

/* shimtest_make_oprofile_protein stripes a protein profile's plain tables the way mf_conversion / vf_conversion / fb_conversion do
 * (src/impl_sse/p7_oprofile.c:773-813, :826-903, :921-1000): rbv 16 bytes per vector (pad 255), rwv 8 words (pad -32768), twv
 * interleaved per stripe with the first four blocks rotated by one node, rfv / tfv 4 floats per vector.
 * iprm = { tbm_b, tec_b, tjb_b, base_b, bias_b, base_w, ddbound_w, xw[E][MOVE], xw[E][LOOP], xw[N][MOVE] }, fprm = { scale_b, scale_w, nj, xfE_move, xfE_loop }. */
P7_OPROFILE *shimtest_make_oprofile_protein(int M, const uint8_t *rbv, const int16_t *rwv, const int16_t *twv, const float *rfv, const float *tfv,
                                            const int *iprm, const float *fprm)
{
  P7_OPROFILE *om = calloc(1, sizeof(P7_OPROFILE));
  const int Kp = 29, Q16 = p7O_NQB(M), Q8 = p7O_NQW(M), Q4 = p7O_NQF(M);
  union { __m128i v; uint8_t b[16]; int16_t w[8]; } u;
  union { __m128 v; float f[4]; } uf;
  int x, q, z, t, j;
  om->M = M; om->allocM = M; om->allocQ16 = Q16; om->allocQ8 = Q8; om->allocQ4 = Q4; om->abc = &shimtest_amino; om->mode = 1;
  om->tbm_b = (uint8_t) iprm[0]; om->tec_b = (uint8_t) iprm[1]; om->tjb_b = (uint8_t) iprm[2]; om->base_b = (uint8_t) iprm[3]; om->bias_b = (uint8_t) iprm[4];
  om->base_w = (int16_t) iprm[5]; om->ddbound_w = (int16_t) iprm[6];
  om->xw[p7O_E][p7O_MOVE] = (int16_t) iprm[7]; om->xw[p7O_E][p7O_LOOP] = (int16_t) iprm[8]; om->xw[p7O_N][p7O_MOVE] = (int16_t) iprm[9];
  om->scale_b = fprm[0]; om->scale_w = fprm[1]; om->nj = fprm[2]; om->xf[p7O_E][p7O_MOVE] = fprm[3]; om->xf[p7O_E][p7O_LOOP] = fprm[4];
  om->rbv_mem = aligned_alloc(16, sizeof(__m128i) * (size_t) Kp * Q16);
  om->rwv_mem = aligned_alloc(16, sizeof(__m128i) * (size_t) Kp * Q8);
  om->twv_mem = aligned_alloc(16, sizeof(__m128i) * (size_t) 8 * Q8);
  om->rfv_mem = aligned_alloc(16, sizeof(__m128) * (size_t) Kp * Q4);
  om->tfv_mem = aligned_alloc(16, sizeof(__m128) * (size_t) 8 * Q4);
  om->rbv = malloc(sizeof(__m128i *) * Kp); om->rwv = malloc(sizeof(__m128i *) * Kp); om->rfv = malloc(sizeof(__m128 *) * Kp);
  om->twv = om->twv_mem; om->tfv = om->tfv_mem;
  for (x = 0; x < Kp; x++) {
    om->rbv[x] = om->rbv_mem + (size_t) x * Q16; om->rwv[x] = om->rwv_mem + (size_t) x * Q8; om->rfv[x] = om->rfv_mem + (size_t) x * Q4;
    for (q = 0; q < Q16; q++) { for (z = 0; z < 16; z++) { const int k = q + 1 + z * Q16; u.b[z] = (k <= M) ? rbv[(size_t) x * (M + 1) + k] : 255; } om->rbv[x][q] = u.v; }
    for (q = 0; q < Q8;  q++) { for (z = 0; z < 8;  z++) { const int k = q + 1 + z * Q8;  u.w[z] = (k <= M) ? rwv[(size_t) x * (M + 1) + k] : -32768; } om->rwv[x][q] = u.v; }
    for (q = 0; q < Q4;  q++) { for (z = 0; z < 4;  z++) { const int k = q + 1 + z * Q4;  uf.f[z] = (k <= M) ? rfv[(size_t) x * (M + 1) + k] : 0.0f; } om->rfv[x][q] = uf.v; }
  }
  for (j = 0, q = 0; q < Q8; q++)
    for (t = p7O_BM; t <= p7O_II; t++) {
      const int kb = (t <= p7O_DM) ? q : q + 1;
      for (z = 0; z < 8; z++) u.w[z] = (kb + z * Q8 < M) ? twv[(size_t) t * (M + 1) + kb + z * Q8] : -32768;
      om->twv[j++] = u.v;
    }
  for (q = 0; q < Q8; q++) { for (z = 0; z < 8; z++) u.w[z] = (q + 1 + z * Q8 < M) ? twv[(size_t) p7O_DD * (M + 1) + q + 1 + z * Q8] : -32768; om->twv[j++] = u.v; }
  for (j = 0, q = 0; q < Q4; q++)
    for (t = p7O_BM; t <= p7O_II; t++) {
      const int kb = (t <= p7O_DM) ? q : q + 1;
      for (z = 0; z < 4; z++) uf.f[z] = (kb + z * Q4 < M) ? tfv[(size_t) t * (M + 1) + kb + z * Q4] : 0.0f;
      om->tfv[j++] = uf.v;
    }
  for (q = 0; q < Q4; q++) { for (z = 0; z < 4; z++) uf.f[z] = (q + 1 + z * Q4 < M) ? tfv[(size_t) p7O_DD * (M + 1) + q + 1 + z * Q4] : 0.0f; om->tfv[j++] = uf.v; }
  return om;
}

void shimtest_free_oprofile_protein(P7_OPROFILE *om)
{
  if (!om) return;
  free(om->rbv_mem); free(om->rwv_mem); free(om->twv_mem); free(om->rfv_mem); free(om->tfv_mem); free(om->rbv); free(om->rwv); free(om->rfv); free(om);
}
void shimtest_set_orf_length(P7_OPROFILE *om, int tjb_b, int xw_move) { om->tjb_b = (uint8_t) tjb_b; om->xw[p7O_N][p7O_MOVE] = (int16_t) xw_move; }
