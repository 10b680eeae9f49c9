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

static const P7_OPROFILE    *shim_loaded_om;    /* the protein profile whose integer tables (and float amino rows) are on the device */

void bathshim_release(void)
{
  if (shim_ctx) bathgpu_destroy(shim_ctx);
  shim_ctx = NULL; shim_loaded = NULL; shim_loaded_om = NULL;
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
    shim_loaded = om_fs; shim_loaded_om = NULL;
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

/* ---- the protein profile's entry points (src/impl_sse/impl_sse.h:488, :518, :544): p7_MSVFilter, p7_ViterbiFilter, p7_ForwardParser ----
 * bathshim_unstripe_oprofile is the un-striping INTEGRATION.md section 2 describes for P7_OPROFILE: rbv[x][q] byte z holds node
 * k = q + Q16 z + 1 (mf_conversion, src/impl_sse/p7_oprofile.c:773-813), rwv[x][q] word z node k = q + Q8 z + 1 and twv the interleaved
 * {BM,MM,IM,DM,MD,MI,II} blocks + Q8 vectors of DD, the first four rotated by one node (vf_conversion, :826-903), rfv / tfv as in
 * P7_FS_OPROFILE.  rbv_out [Kp][M+1] uint8, rwv_out [Kp][M+1] int16, twv_out [8][M+1] int16 (source-node indexed), rfv_out [Kp][M+1],
 * tfv_out [8][M+1]. */
void bathshim_unstripe_oprofile(const P7_OPROFILE *om, uint8_t *rbv_out, int16_t *rwv_out, int16_t *twv_out, float *rfv_out, float *tfv_out)
{
  const int M = om->M, Kp = om->abc->Kp, Q16 = p7O_NQB(M), Q8 = p7O_NQW(M), Q4 = p7O_NQF(M);
  union { __m128i v; uint8_t b[16]; int16_t w[8]; } u;
  union { __m128 v; float f[4]; } uf;
  int x, q, z, t, k;
  memset(rbv_out, 255, (size_t) Kp * (M + 1));
  for (x = 0; x < Kp; x++)
    for (q = 0; q < Q16; q++) { u.v = om->rbv[x][q]; for (z = 0; z < 16; z++) { k = q + Q16 * z + 1; if (k <= M) rbv_out[(size_t) x * (M + 1) + k] = u.b[z]; } }
  for (k = 0; k < Kp * (M + 1); k++) rwv_out[k] = -32768;
  for (x = 0; x < Kp; x++)
    for (q = 0; q < Q8; q++) { u.v = om->rwv[x][q]; for (z = 0; z < 8; z++) { k = q + Q8 * z + 1; if (k <= M) rwv_out[(size_t) x * (M + 1) + k] = u.w[z]; } }
  for (k = 0; k < 8 * (M + 1); k++) twv_out[k] = -32768;
  for (q = 0; q < Q8; q++)
    for (t = p7O_BM; t <= p7O_II; t++) {
      u.v = om->twv[7 * q + t];
      for (z = 0; z < 8; z++) {
        k = q + Q8 * z + 1;
        if (k > M) continue;
        if (t <= p7O_DM) twv_out[(size_t) t * (M + 1) + (k - 1)] = u.w[z];
        else             twv_out[(size_t) t * (M + 1) + k]       = u.w[z];
      }
    }
  for (q = 0; q < Q8; q++) { u.v = om->twv[7 * Q8 + q]; for (z = 0; z < 8; z++) { k = q + Q8 * z + 1; if (k <= M) twv_out[(size_t) p7O_DD * (M + 1) + k] = u.w[z]; } }
  memset(rfv_out, 0, sizeof(float) * (size_t) Kp * (M + 1));
  memset(tfv_out, 0, sizeof(float) * (size_t) 8 * (M + 1));
  for (x = 0; x < Kp; x++)
    for (q = 0; q < Q4; q++) { uf.v = om->rfv[x][q]; for (z = 0; z < 4; z++) { k = q + Q4 * z + 1; if (k <= M) rfv_out[(size_t) x * (M + 1) + k] = uf.f[z]; } }
  for (q = 0; q < Q4; q++)
    for (t = p7O_BM; t <= p7O_II; t++) {
      uf.v = om->tfv[7 * q + t];
      for (z = 0; z < 4; z++) {
        k = q + Q4 * z + 1;
        if (k > M) continue;
        if (t <= p7O_DM) tfv_out[(size_t) t * (M + 1) + (k - 1)] = uf.f[z];
        else             tfv_out[(size_t) t * (M + 1) + k]       = uf.f[z];
      }
    }
  for (q = 0; q < Q4; q++) { uf.v = om->tfv[7 * Q4 + q]; for (z = 0; z < 4; z++) { k = q + Q4 * z + 1; if (k <= M) tfv_out[(size_t) p7O_DD * (M + 1) + k] = uf.f[z]; } }
}

static int shim_prepare_om(const P7_OPROFILE *om)
{
  int status;
  if (!shim_ctx && (status = bathgpu_create(shim_device, &shim_ctx)) != BATHGPU_OK) return status;
  if (shim_loaded_om != om) {
    const int M = om->M, Kp = om->abc->Kp, nrows = p7P_MAXCODONS3 + Kp;
    uint8_t *rbv = malloc((size_t) Kp * (M + 1));
    int16_t *rwv = malloc(sizeof(int16_t) * (size_t) Kp * (M + 1)), *twv = malloc(sizeof(int16_t) * (size_t) 8 * (M + 1));
    float   *rfv = calloc((size_t) nrows * (M + 1), sizeof(float)), *tfv = malloc(sizeof(float) * (size_t) 8 * (M + 1));
    bathgpu_filter_params fp;
    if (!rbv || !rwv || !twv || !rfv || !tfv) { free(rbv); free(rwv); free(twv); free(rfv); free(tfv); return eslEMEM; }
    /* the float kernels read the amino-acid rows of a 3-codon image (rows 338 + x): an image whose codon rows are empty will do */
    bathshim_unstripe_oprofile(om, rbv, rwv, twv, rfv + (size_t) p7P_MAXCODONS3 * (M + 1), tfv);
    fp.M = M; fp.tbm_b = om->tbm_b; fp.tec_b = om->tec_b; fp.base_b = om->base_b; fp.bias_b = om->bias_b; fp.scale_b = om->scale_b;
    fp.base_w = om->base_w; fp.ddbound_w = om->ddbound_w; fp.xw_E_move = om->xw[p7O_E][p7O_MOVE]; fp.xw_E_loop = om->xw[p7O_E][p7O_LOOP];
    fp.scale_w = om->scale_w; fp.cpu_lanes_u8 = 16; fp.cpu_lanes_i16 = 8;
    status = bathgpu_load_filter_profile(shim_ctx, &fp, rbv, rwv, twv);
    if (status == BATHGPU_OK) status = bathgpu_load_fs_profile(shim_ctx, 3, M, nrows, rfv, tfv);
    free(rbv); free(rwv); free(twv); free(rfv); free(tfv);
    if (status != BATHGPU_OK) return status;
    shim_loaded_om = om; shim_loaded = NULL;
  }
  return eslOK;
}

static int shim_one_orf(const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, bathgpu_orf *o)
{
  int status;
  if ((status = shim_prepare_om(om)) != eslOK) return status;
  if ((status = bathgpu_select_slot(shim_ctx, 0)) != BATHGPU_OK) return status;
  if ((status = bathgpu_upload_orfs(shim_ctx, dsq + 1, (int64_t) L)) != BATHGPU_OK) return status;
  memset(o, 0, sizeof *o);
  o->offset = 0; o->L = L;
  o->tjb_b   = om->tjb_b;                          /* what p7_oprofile_ReconfigLength(om, L) left in the profile */
  o->xw_move = om->xw[p7O_N][p7O_MOVE];
  return eslOK;
}

int p7_MSVFilter(const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, P7_OMX *ox, float *ret_sc)
{
  bathgpu_orf o; float sc; int32_t st; int status;
  (void) ox;
  if ((status = shim_one_orf(dsq, L, om, &o)) != eslOK) return status;
  if ((status = bathgpu_msv_orfs(shim_ctx, &o, 1, &sc, &st)) != BATHGPU_OK) return status;
  *ret_sc = sc;
  return st;                                       /* eslOK, or eslERANGE with *ret_sc = +infinity (msvfilter.c:176-180) */
}

int p7_ViterbiFilter(const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, P7_OMX *ox, float *ret_sc)
{
  bathgpu_orf o; float sc; int32_t st; int nw = 0, status;
  (void) ox;
  if ((status = shim_one_orf(dsq, L, om, &o)) != eslOK) return status;
  if ((status = bathgpu_vit_orfs(shim_ctx, &o, 1, &sc, &st, NULL, 0, &nw)) != BATHGPU_OK) return status;
  *ret_sc = sc;
  return st;
}

int p7_ForwardParser(const ESL_DSQ *dsq, int L, const P7_OPROFILE *om, P7_OMX *fwd, float *opt_sc)
{
  bathgpu_orf o; float sc; int32_t st; int status;
  const float xfE[2] = { om->xf[p7O_E][p7O_MOVE], om->xf[p7O_E][p7O_LOOP] };
  (void) fwd;                                      /* scores only: the X rows come from bathgpu_orf_fwd_bck_xrows in the batched wiring */
  if ((status = shim_one_orf(dsq, L, om, &o)) != eslOK) return status;
  if ((status = bathgpu_fwd_orfs(shim_ctx, &o, 1, om->nj, xfE, &sc, &st)) != BATHGPU_OK) return status;
  if (opt_sc) *opt_sc = sc;
  return st;
}
