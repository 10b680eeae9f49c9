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
