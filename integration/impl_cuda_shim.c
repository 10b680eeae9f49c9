/* impl_cuda_shim.c -- the impl-layer entry points of the frameshift parsers, with the reference's own signatures, on libbathgpu.so.
 *
 * BATH's DP kernels sit behind the impl layer (src/hmmer.h:1044-1052; prototypes src/impl_sse/impl_sse.h:483-534).  This file
 * implements the two parser prototypes of src/impl_sse/impl_sse.h:493-494 VERBATIM,
 *
 *   int p7_ForwardParser_Frameshift_3Codons (const ESL_DSQ *dsq, int L, const P7_FS_OPROFILE *om_fs, P7_OMX *ox, P7_OIVX *ov, float *opt_sc);
 *   int p7_BackwardParser_Frameshift_3Codons(const ESL_DSQ *dsq, int L, const P7_FS_OPROFILE *om_fs, const P7_OMX *fwd, P7_OMX *bck,
 *                                            P7_OIVX *ov, float *opt_sc);
 *
 * as batch-of-one calls into the batched C ABI of include/bathgpu.h, so that a BATH tree can link this object in place of
 * src/impl_sse/fwdback_fs.c's two functions without touching a caller: same arguments, same fields of P7_OMX written (M, L,
 * has_own_scales, totscale, the xmx rows {E,N,J,B,C,SCALE} for i = 0..L: fwdback_fs.c:132-135,159-165,305-309,492-505), same return
 * codes (eslOK / eslERANGE with *opt_sc = -inf, :521-526).  One window per call cannot feed a GPU -- the batched glue of
 * INTEGRATION.md section 2 is the production wiring; this shim exists so that the boundary is CHECKABLE: it compiles against the
 * reference's struct layout (bath_impl_standin.h here, impl_sse.h in a BATH tree), and tests/test_impl_shim.py stripes a profile the
 * way fs_fb_conversion does (src/impl_sse/p7_fs_oprofile.c:222-296), pushes it through bathshim_unstripe_fs_profile and through these
 * functions, and compares with the oracle.
 *
 * bathshim_unstripe_fs_profile is the un-striping INTEGRATION.md section 1 describes: rfv[c][q] lane z holds node k = q + Q z + 1;
 * tfv holds per stripe q the vectors {BM,MM,IM,DM,MD,MI,II}, the first four rotated by one node (source node k-1), then Q vectors of DD.
 */
#ifdef BATH_SHIM_USE_REFERENCE_HEADERS
#include "hmmer.h"
#include "impl_sse.h"
#else
#include <math.h>
#include "bath_impl_standin.h"
#endif
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "bathgpu.h"

static bathgpu_ctx          *shim_ctx;          /* one context per process: BATH's worker threads each clone a pipeline (src/bathsearch.c:814-844); a threaded build keeps one per worker */
static const P7_FS_OPROFILE *shim_loaded;       /* the profile whose image is on the device */
static int                   shim_device = 0;

void bathshim_set_device(int device) { shim_device = device; }

void bathshim_release(void)
{
  if (shim_ctx) bathgpu_destroy(shim_ctx);
  shim_ctx = NULL; shim_loaded = NULL;
}

/* rfv_out [nrows][M+1], tfv_out [8][M+1] (BM MM IM DM MD MI II DD, indexed by SOURCE node), both zero-filled by the caller or not:
 * every entry is written.  Returns the number of emission rows. */
int bathshim_unstripe_fs_profile(const P7_FS_OPROFILE *om, float *rfv_out, float *tfv_out)
{
  const int M = om->M, Q = p7O_NQF(M);
  const int nrows = (om->codon_lengths == 3 ? p7P_MAXCODONS3 : p7P_MAXCODONS5) + om->abc->Kp;
  union { __m128 v; float f[4]; } u;
  int c, q, z, t, k;
  memset(rfv_out, 0, sizeof(float) * (size_t) nrows * (size_t) (M + 1));
  memset(tfv_out, 0, sizeof(float) * (size_t) p7O_NTRANS * (size_t) (M + 1));
  for (c = 0; c < nrows; c++)
    for (q = 0; q < Q; q++) {
      u.v = om->rfv[c][q];
      for (z = 0; z < 4; z++) { k = q + Q * z + 1; if (k <= M) rfv_out[(size_t) c * (M + 1) + k] = u.f[z]; }
    }
  for (q = 0; q < Q; q++)
    for (t = p7O_BM; t <= p7O_II; t++) {
      u.v = om->tfv[7 * q + t];
      for (z = 0; z < 4; z++) {
        k = q + Q * z + 1;
        if (k > M) continue;
        if (t <= p7O_DM) tfv_out[(size_t) t * (M + 1) + (k - 1)] = u.f[z];      /* rotated: the transition out of node k-1 */
        else             tfv_out[(size_t) t * (M + 1) + k]       = u.f[z];
      }
    }
  for (q = 0; q < Q; q++) {
    u.v = om->tfv[7 * Q + q];
    for (z = 0; z < 4; z++) { k = q + Q * z + 1; if (k <= M) tfv_out[(size_t) p7O_DD * (M + 1) + k] = u.f[z]; }
  }
  return nrows;
}

static int shim_prepare(const P7_FS_OPROFILE *om_fs)
{
  int status;
  if (!shim_ctx && (status = bathgpu_create(shim_device, &shim_ctx)) != BATHGPU_OK) return status;
  if (shim_loaded != om_fs) {
    const int M = om_fs->M;
    const int nrows = (om_fs->codon_lengths == 3 ? p7P_MAXCODONS3 : p7P_MAXCODONS5) + om_fs->abc->Kp;
    float *rfv = malloc(sizeof(float) * (size_t) nrows * (size_t) (M + 1));
    float *tfv = malloc(sizeof(float) * (size_t) p7O_NTRANS * (size_t) (M + 1));
    if (!rfv || !tfv) { free(rfv); free(tfv); return eslEMEM; }
    bathshim_unstripe_fs_profile(om_fs, rfv, tfv);
    status = bathgpu_load_fs_profile(shim_ctx, om_fs->codon_lengths, M, nrows, rfv, tfv);
    free(rfv); free(tfv);
    if (status != BATHGPU_OK) return status;
    shim_loaded = om_fs;
  }
  return eslOK;
}

/* the window descriptor of a whole sequence under the profile's CURRENT length model (what p7_fs_oprofile_ReconfigLength left in xf) */
static bathgpu_window whole_window(const P7_FS_OPROFILE *om_fs, int L)
{
  bathgpu_window w;
  w.start = 1; w.L = L; w.pmove = om_fs->xf[p7O_N][p7O_MOVE]; w.ploop = om_fs->xf[p7O_N][p7O_LOOP];
  return w;
}

static int run_parsers(const ESL_DSQ *dsq, int L, const P7_FS_OPROFILE *om_fs, float *fx, float *bx, float *fsc, float *bsc, int32_t *st)
{
  bathgpu_window w = whole_window(om_fs, L);
  const float xfE[2] = { om_fs->xf[p7O_E][p7O_MOVE], om_fs->xf[p7O_E][p7O_LOOP] };
  int status;
  if (om_fs->codon_lengths != 3) return eslEINVAL;                       /* fwdback_fs.c:123 */
  if ((status = shim_prepare(om_fs)) != eslOK) return status;
  if ((status = bathgpu_select_slot(shim_ctx, 0)) != BATHGPU_OK) return status;
  if ((status = bathgpu_upload_block(shim_ctx, dsq, (int64_t) L)) != BATHGPU_OK) return status;
  return bathgpu_fs_fwd_bck_xrows(shim_ctx, &w, 1, xfE, fx, bx, fsc, bsc, st);
}

static void fill_omx(P7_OMX *ox, const P7_FS_OPROFILE *om_fs, int L, const float *xr)
{
  int i;
  ox->M = om_fs->M; ox->L = L; ox->has_own_scales = 1; ox->totscale = 0.0;
  memcpy(ox->xmx, xr, sizeof(float) * (size_t) (L + 1) * p7X_NXCELLS);
  for (i = 0; i <= L; i++) if (xr[i * p7X_NXCELLS + p7X_SCALE] != 1.0f) ox->totscale += log(xr[i * p7X_NXCELLS + p7X_SCALE]);
}

int p7_ForwardParser_Frameshift_3Codons(const ESL_DSQ *dsq, int L, const P7_FS_OPROFILE *om_fs, P7_OMX *ox, P7_OIVX *ov, float *opt_sc)
{
  float *fx = malloc(sizeof(float) * (size_t) (L + 1) * p7X_NXCELLS * 2), *bx = fx ? fx + (size_t) (L + 1) * p7X_NXCELLS : NULL;
  float fsc, bsc; int32_t st; int status;
  (void) ov;
  if (!fx) return eslEMEM;
  status = run_parsers(dsq, L, om_fs, fx, bx, &fsc, &bsc, &st);
  if (status == eslOK) {
    fill_omx(ox, om_fs, L, fx);
    if (opt_sc) *opt_sc = fsc;
    status = st;                                                         /* eslERANGE where the reference returns it (:521-526) */
  }
  free(fx);
  return status;
}

int p7_BackwardParser_Frameshift_3Codons(const ESL_DSQ *dsq, int L, const P7_FS_OPROFILE *om_fs, const P7_OMX *fwd, P7_OMX *bck, P7_OIVX *ov,
                                         float *opt_sc)
{
  float *fx = malloc(sizeof(float) * (size_t) (L + 1) * p7X_NXCELLS * 2), *bx = fx ? fx + (size_t) (L + 1) * p7X_NXCELLS : NULL;
  float fsc, bsc; int32_t st; int status;
  (void) ov; (void) fwd;                                                 /* the device re-runs Forward for its scale factors: fwd must be this sequence's */
  if (!fx) return eslEMEM;
  status = run_parsers(dsq, L, om_fs, fx, bx, &fsc, &bsc, &st);
  if (status == eslOK) {
    fill_omx(bck, om_fs, L, bx);
    bck->has_own_scales = 0;                                             /* fwdback_fs.c:609: Backward runs on Forward's scale factors */
    if (opt_sc) *opt_sc = bsc;
    status = st;
  }
  free(fx);
  return status;
}
