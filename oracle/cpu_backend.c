/* cpu_backend.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 *
 * The batched stage calls of include/bathgpu.h implemented on the CPU with the oracle's scalar functions and a
 * pool of POSIX threads -- the way the reference runs the same stages on its worker threads.  It exists so that
 *   (1) the host pipeline (bath_b200/host/pipeline.cpp) can be checked against the reference's golden outputs on a
 *       machine without a GPU (tests/test_host_pipeline_cpu.py), and
 *   (2) bench.py's cpu_baseline / --impl reference legs can time the whole translated search on the host cores.
 * It is handed to the host pipeline as a bathhost_backend table by TESTS AND BENCH ONLY; the product never loads it.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <pthread.h>
#include "bath_oracle.h"
#include "../include/bathgpu.h"

typedef struct {
  int             nthreads;
  char            err[256];
  BO_FS_OPROFILE *om3, *om5;
  BO_OPROFILE    *om;             /* integer tables from load_filter_profile; float tables from the 3-codon image */
  uint8_t        *ssv_scores;
  int             lanes_u8, lanes_i16;
  int             cur, nslots;    /* target slots grow on demand (bathgpu_select_slot) */
  uint8_t       **dsq;      int64_t *n;
  uint8_t       **res;      int64_t *nres;
  bathgpu_orf_hit **hits;   int64_t *nhits;     /* per slot: survivors of the last screen */
  int             use_simd;       /* bo_backend_set_simd: the MSV screen on the AVX2 build (msv_avx2.c; bench.py's CPU arm) */
  bo_msv_simd    *msv_simd;
} bo_backend;

static int grow_slots(bo_backend *b, int want)
{
  int s;
  if (want <= b->nslots) return BO_OK;
  b->dsq   = realloc(b->dsq,   sizeof(uint8_t *) * (size_t) want);          b->n     = realloc(b->n,     sizeof(int64_t) * (size_t) want);
  b->res   = realloc(b->res,   sizeof(uint8_t *) * (size_t) want);          b->nres  = realloc(b->nres,  sizeof(int64_t) * (size_t) want);
  b->hits  = realloc(b->hits,  sizeof(bathgpu_orf_hit *) * (size_t) want);  b->nhits = realloc(b->nhits, sizeof(int64_t) * (size_t) want);
  if (!b->dsq || !b->n || !b->res || !b->nres || !b->hits || !b->nhits) return BO_EMEM;
  for (s = b->nslots; s < want; s++) { b->dsq[s] = NULL; b->n[s] = 0; b->res[s] = NULL; b->nres[s] = 0; b->hits[s] = NULL; b->nhits[s] = 0; }
  b->nslots = want;
  return BO_OK;
}

bo_backend *bo_backend_create(int nthreads)
{
  bo_backend *b = calloc(1, sizeof(bo_backend));
  if (b) { b->nthreads = nthreads > 0 ? nthreads : 1; grow_slots(b, 2); }
  bo_FLogsumInit();
  return b;
}

void bo_backend_destroy(bo_backend *b)
{
  int s;
  if (!b) return;
  bo_fs_oprofile_destroy(b->om3); bo_fs_oprofile_destroy(b->om5); bo_oprofile_destroy(b->om);
  free(b->ssv_scores);
  bo_msv_simd_destroy(b->msv_simd);
  for (s = 0; s < b->nslots; s++) { free(b->dsq[s]); free(b->res[s]); free(b->hits[s]); }
  free(b->dsq); free(b->n); free(b->res); free(b->nres); free(b->hits); free(b->nhits);
  free(b);
}

/* The MSV + SSV screen of every ORF on the AVX2 build of the same byte arithmetic (returns 1 if it is in use).  The scalar restatement
 * stays the default: tests check the device against it; bench.py's CPU arm asks for the SIMD one. */
int bo_backend_set_simd(bo_backend *b, int on)
{
  if (!b) return 0;
  b->use_simd = on && bo_msv_simd_supported();
  if (b->use_simd && b->om && !b->msv_simd) b->msv_simd = bo_msv_simd_create(b->om);
  return b->use_simd;
}

const char *bo_backend_last_error(const void *ctx) { return ctx ? ((const bo_backend *) ctx)->err : "no context"; }

/* ---- parallel for over items ---- */
typedef void (*item_fn)(bo_backend *b, void *arg, int item);
typedef struct { bo_backend *b; void *arg; item_fn fn; int n, next; pthread_mutex_t mu; } pf_job;
static void *pf_worker(void *p)
{
  pf_job *j = p;
  for (;;) {
    int i0, i1, i;
    pthread_mutex_lock(&j->mu); i0 = j->next; j->next += 16; pthread_mutex_unlock(&j->mu);
    if (i0 >= j->n) break;
    i1 = i0 + 16 < j->n ? i0 + 16 : j->n;
    for (i = i0; i < i1; i++) j->fn(j->b, j->arg, i);
  }
  return NULL;
}
static void parallel_for(bo_backend *b, int n, item_fn fn, void *arg)
{
  pf_job j; pthread_t th[64]; int t, nt = b->nthreads > 64 ? 64 : b->nthreads;
  if (nt > n) nt = n > 0 ? n : 1;
  j.b = b; j.arg = arg; j.fn = fn; j.n = n; j.next = 0; pthread_mutex_init(&j.mu, NULL);
  for (t = 0; t < nt; t++) pthread_create(&th[t], NULL, pf_worker, &j);
  for (t = 0; t < nt; t++) pthread_join(th[t], NULL);
  pthread_mutex_destroy(&j.mu);
}

/* ---- profiles ---- */
int bo_backend_load_fs_profile(void *ctx, int which, int M, int nrows, const float *rfv, const float *tfv)
{
  bo_backend *b = ctx;
  BO_FS_OPROFILE *om = calloc(1, sizeof(BO_FS_OPROFILE));
  if (!om) return BO_EMEM;
  om->M = M; om->L = 100; om->mode = BO_LOCAL; om->codon_lengths = which; om->nrows = nrows; om->maxcodons = nrows - BO_KP; om->nj = 1.0f;
  om->rfv = malloc(sizeof(float) * (size_t) nrows * (M + 1));
  om->tfv = malloc(sizeof(float) * (size_t) 8 * (M + 1));
  memcpy(om->rfv, rfv, sizeof(float) * (size_t) nrows * (M + 1));
  memcpy(om->tfv, tfv, sizeof(float) * (size_t) 8 * (M + 1));
  om->xf[BO_X_E][BO_O_MOVE] = 0.5f; om->xf[BO_X_E][BO_O_LOOP] = 0.5f;
  if (which == 3) { bo_fs_oprofile_destroy(b->om3); b->om3 = om; } else { bo_fs_oprofile_destroy(b->om5); b->om5 = om; }
  return BO_OK;
}

int bo_backend_load_filter_profile(void *ctx, const void *prm_, const uint8_t *rbv, const int16_t *rwv, const int16_t *twv)
{
  bo_backend *b = ctx;
  const bathgpu_filter_params *p = prm_;
  int M = p->M, x, k;
  BO_OPROFILE *om;
  if (!b->om3 || b->om3->M != M) { snprintf(b->err, sizeof b->err, "load the 3-codon profile first"); return BO_EINVAL; }
  om = calloc(1, sizeof(BO_OPROFILE));
  om->M = M; om->L = 100; om->mode = BO_LOCAL; om->nj = 1.0f;
  om->tbm_b = p->tbm_b; om->tec_b = p->tec_b; om->base_b = p->base_b; om->bias_b = p->bias_b; om->scale_b = p->scale_b;
  om->base_w = p->base_w; om->ddbound_w = p->ddbound_w; om->scale_w = p->scale_w;
  om->xw[BO_X_E][BO_O_MOVE] = p->xw_E_move; om->xw[BO_X_E][BO_O_LOOP] = p->xw_E_loop;
  om->rbv = malloc((size_t) BO_KP * (M + 1));                       memcpy(om->rbv, rbv, (size_t) BO_KP * (M + 1));
  om->rwv = malloc(sizeof(int16_t) * (size_t) BO_KP * (M + 1));     memcpy(om->rwv, rwv, sizeof(int16_t) * (size_t) BO_KP * (M + 1));
  om->twv = malloc(sizeof(int16_t) * (size_t) 8 * (M + 1));         memcpy(om->twv, twv, sizeof(int16_t) * (size_t) 8 * (M + 1));
  /* float tables: the amino rows and transitions of the 3-codon image are the protein profile's (modelconfig.c:343-352) */
  om->rfv = malloc(sizeof(float) * (size_t) BO_KP * (M + 1));
  om->tfv = malloc(sizeof(float) * (size_t) 8 * (M + 1));
  for (x = 0; x < BO_KP; x++) for (k = 0; k <= M; k++) om->rfv[(size_t) x * (M + 1) + k] = b->om3->rfv[(size_t)(b->om3->maxcodons + x) * (M + 1) + k];
  memcpy(om->tfv, b->om3->tfv, sizeof(float) * (size_t) 8 * (M + 1));
  om->xf[BO_X_E][BO_O_MOVE] = 0.5f; om->xf[BO_X_E][BO_O_LOOP] = 0.5f;
  bo_oprofile_destroy(b->om); b->om = om;
  free(b->ssv_scores);
  b->ssv_scores = malloc((size_t)(M + 1) * BO_KP);
  bo_oprofile_ssv_scores(om, b->ssv_scores);
  b->lanes_u8 = p->cpu_lanes_u8; b->lanes_i16 = p->cpu_lanes_i16;
  bo_msv_simd_destroy(b->msv_simd); b->msv_simd = NULL;
  if (b->use_simd) b->msv_simd = bo_msv_simd_create(om);
  return BO_OK;
}

/* ---- targets ---- */
int bo_backend_select_slot(void *ctx, int slot)
{
  bo_backend *b = ctx;
  if (slot < 0 || slot >= 65536 || grow_slots(b, slot + 1) != BO_OK) return BO_EINVAL;
  b->cur = slot;
  return BO_OK;
}

int bo_backend_upload_block(void *ctx, const uint8_t *dsq, int64_t n)
{
  bo_backend *b = ctx;
  free(b->dsq[b->cur]);
  b->dsq[b->cur] = malloc((size_t) n + 2);
  memcpy(b->dsq[b->cur], dsq, (size_t) n + 2);
  b->n[b->cur] = n;
  return BO_OK;
}

/* bathgpu_upload_block_segments: the block handed over as pieces (seg[g] points at the first nucleotide of piece g) */
int bo_backend_upload_block_segments(void *ctx, const uint8_t *const *seg, const int64_t *seg_n, int nseg)
{
  bo_backend *b = ctx;
  int64_t n = 0, off = 0;
  int g;
  if (!seg || !seg_n || nseg < 1) return BO_EINVAL;
  for (g = 0; g < nseg; g++) { if (!seg[g] || seg_n[g] < 1) return BO_EINVAL; n += seg_n[g]; }
  free(b->dsq[b->cur]);
  b->dsq[b->cur] = malloc((size_t) n + 2);
  b->dsq[b->cur][0] = 255; b->dsq[b->cur][n + 1] = 255;
  for (g = 0; g < nseg; g++) { memcpy(b->dsq[b->cur] + 1 + off, seg[g], (size_t) seg_n[g]); off += seg_n[g]; }
  b->n[b->cur] = n;
  return BO_OK;
}

int bo_backend_revcomp_slot(void *ctx, int src, int dst)
{
  bo_backend *b = ctx;
  int64_t n;
  if (src == dst || src < 0 || dst < 0 || src >= 65536 || dst >= 65536 || grow_slots(b, (src > dst ? src : dst) + 1) != BO_OK || !b->dsq[src]) return BO_EINVAL;
  n = b->n[src];
  free(b->dsq[dst]);
  b->dsq[dst] = malloc((size_t) n + 2);
  memcpy(b->dsq[dst], b->dsq[src], (size_t) n + 2);
  bo_dna_revcomp(b->dsq[dst], n);
  b->n[dst] = n;
  return BO_OK;
}

int bo_backend_upload_orfs(void *ctx, const uint8_t *residues, int64_t n)
{
  bo_backend *b = ctx;
  free(b->res[b->cur]);
  b->res[b->cur] = malloc((size_t) n);
  memcpy(b->res[b->cur], residues, (size_t) n);
  b->nres[b->cur] = n;
  return BO_OK;
}

static uint8_t *orf_dsq(const bo_backend *b, const bathgpu_orf *o)
{
  uint8_t *d = malloc((size_t) o->L + 2);
  d[0] = d[o->L + 1] = BO_DSQ_SENTINEL;
  memcpy(d + 1, b->res[b->cur] + o->offset, (size_t) o->L);
  return d;
}

/* ---- ORF stages ---- */
typedef struct { const bathgpu_orf *orfs; float *sc; int32_t *st; BO_WINDOWLIST *wl; float nj; const float *xfE; int mode; } orf_args;

static void orf_item(bo_backend *b, void *arg, int i)
{
  orf_args *a = arg;
  const bathgpu_orf *o = &a->orfs[i];
  BO_OPROFILE om = *b->om;                      /* private per-length pieces; tables shared */
  uint8_t *d = orf_dsq(b, o);
  om.tjb_b = o->tjb_b;
  om.xw[BO_X_N][BO_O_MOVE] = om.xw[BO_X_C][BO_O_MOVE] = om.xw[BO_X_J][BO_O_MOVE] = o->xw_move;
  om.xw[BO_X_N][BO_O_LOOP] = om.xw[BO_X_C][BO_O_LOOP] = om.xw[BO_X_J][BO_O_LOOP] = 0;
  switch (a->mode) {
  case 0: a->st[i] = bo_MSVFilter(d, o->L, &om, &a->sc[i]); break;
  case 1: bo_SSVFilter_BATH_thresh(d, o->L, &om, b->ssv_scores, o->ssv_thresh, b->lanes_u8, &a->wl[i]); break;
  case 2: a->st[i] = bo_ViterbiFilter_BATH_thresh(d, o->L, &om, b->ssv_scores, o->vit_thresh, o->ext_thresh, b->lanes_i16,
                                                  (o->flags & 1) ? &a->wl[i] : NULL, &a->sc[i]); break;
  case 3: {
    float pmove = (2.0f + a->nj) / ((float) o->L + 2.0f + a->nj), ploop = 1.0f - pmove;
    om.xf[BO_X_N][BO_O_LOOP] = om.xf[BO_X_C][BO_O_LOOP] = om.xf[BO_X_J][BO_O_LOOP] = ploop;
    om.xf[BO_X_N][BO_O_MOVE] = om.xf[BO_X_C][BO_O_MOVE] = om.xf[BO_X_J][BO_O_MOVE] = pmove;
    om.xf[BO_X_E][BO_O_MOVE] = a->xfE[0]; om.xf[BO_X_E][BO_O_LOOP] = a->xfE[1];
    a->sc[i] = 0.0f;
    a->st[i] = bo_ForwardParser(d, o->L, &om, &a->sc[i]);
    break; }
  }
  free(d);
}

static int collect_windows(BO_WINDOWLIST *wl, int n, bathgpu_orf_window *out, int max_wins, int *nwins)
{
  int i, z, tot = 0, rc = BO_OK;
  for (i = 0; i < n; i++) {
    for (z = 0; z < wl[i].count; z++) {
      if (tot < max_wins) { out[tot].orf = i; out[tot].n = wl[i].w[z].n; out[tot].k = wl[i].w[z].k; out[tot].length = wl[i].w[z].length; out[tot].score = wl[i].w[z].score; }
      else rc = BO_EINVAL;
      tot++;
    }
    bo_windowlist_free(&wl[i]);
  }
  *nwins = tot;
  return rc;
}

int bo_backend_msv_orfs(void *ctx, const void *orfs, int n, float *sc, int32_t *status)
{
  orf_args a = { orfs, sc, status, NULL, 0, NULL, 0 };
  parallel_for(ctx, n, orf_item, &a);
  return BO_OK;
}

int bo_backend_ssv_windows(void *ctx, const void *orfs, int n, void *wins, int max_wins, int *nwins)
{
  BO_WINDOWLIST *wl = calloc((size_t) n, sizeof(BO_WINDOWLIST));
  orf_args a = { orfs, NULL, NULL, wl, 0, NULL, 1 };
  int rc;
  parallel_for(ctx, n, orf_item, &a);
  rc = collect_windows(wl, n, wins, max_wins, nwins);
  free(wl);
  return rc;
}

int bo_backend_vit_orfs(void *ctx, const void *orfs, int n, float *sc, int32_t *status, void *wins, int max_wins, int *nwins)
{
  BO_WINDOWLIST *wl = calloc((size_t) n, sizeof(BO_WINDOWLIST));
  orf_args a = { orfs, sc, status, wl, 0, NULL, 2 };
  int rc, nw = 0;
  parallel_for(ctx, n, orf_item, &a);
  rc = collect_windows(wl, n, wins, wins ? max_wins : 0, &nw);
  if (nwins) *nwins = nw;
  free(wl);
  return (wins || nw == 0) ? rc : BO_OK;
}

int bo_backend_fwd_orfs(void *ctx, const void *orfs, int n, float nj, const float xfE[2], float *fwdsc, int32_t *status)
{
  orf_args a = { orfs, fwdsc, status, NULL, nj, xfE, 3 };
  parallel_for(ctx, n, orf_item, &a);
  return BO_OK;
}

/* ---- window stages ---- */
typedef struct { const bathgpu_window *w; const float *xfE; float *fsc, *bsc; int32_t *st; float *fx, *bx; const int64_t *xoff; int do_bck; } win_args;

static void win_item(bo_backend *b, void *arg, int i)
{
  win_args *a = arg;
  const bathgpu_window *w = &a->w[i];
  BO_FS_OPROFILE om = *b->om3;
  int L = w->L, st;
  uint8_t *sub = malloc((size_t) L + 2);
  BO_MX *oxf = bo_mx_create(om.M, L, 0), *oxb = NULL;
  float fsc = 0.0f, bsc = 0.0f;
  sub[0] = sub[L + 1] = BO_DSQ_SENTINEL;
  memcpy(sub + 1, b->dsq[b->cur] + w->start, (size_t) L);
  om.xf[BO_X_N][BO_O_LOOP] = om.xf[BO_X_C][BO_O_LOOP] = om.xf[BO_X_J][BO_O_LOOP] = w->ploop;
  om.xf[BO_X_N][BO_O_MOVE] = om.xf[BO_X_C][BO_O_MOVE] = om.xf[BO_X_J][BO_O_MOVE] = w->pmove;
  om.xf[BO_X_E][BO_O_MOVE] = a->xfE[0]; om.xf[BO_X_E][BO_O_LOOP] = a->xfE[1];
  st = bo_ForwardParser_Frameshift_3Codons(sub, L, &om, oxf, &fsc);
  if (a->do_bck && st == BO_OK) {
    oxb = bo_mx_create(om.M, L, 0);
    st = bo_BackwardParser_Frameshift_3Codons(sub, L, &om, oxf, oxb, &bsc);
    memcpy(a->fx + a->xoff[i] * 6, oxf->xmx, sizeof(float) * (size_t)(L + 1) * 6);
    memcpy(a->bx + a->xoff[i] * 6, oxb->xmx, sizeof(float) * (size_t)(L + 1) * 6);
  }
  if (a->fsc) a->fsc[i] = fsc;
  if (a->bsc) a->bsc[i] = bsc;
  a->st[i] = st;
  bo_mx_destroy(oxf); bo_mx_destroy(oxb); free(sub);
}

int bo_backend_fs_fwd_windows(void *ctx, const void *wins, int n, const float xfE[2], float *fwdsc, int32_t *status)
{
  win_args a = { wins, xfE, fwdsc, NULL, status, NULL, NULL, NULL, 0 };
  parallel_for(ctx, n, win_item, &a);
  return BO_OK;
}

int bo_backend_fs_fwd_bck_xrows(void *ctx, const void *wins, int n, const float xfE[2], float *fwd_xrows, float *bck_xrows,
                                float *fwdsc, float *bcksc, int32_t *status)
{
  const bathgpu_window *w = wins;
  int64_t *xoff = malloc(sizeof(int64_t) * (size_t)(n + 1));
  int i;
  win_args a = { wins, xfE, fwdsc, bcksc, status, fwd_xrows, bck_xrows, xoff, 1 };
  xoff[0] = 0;
  for (i = 0; i < n; i++) xoff[i + 1] = xoff[i] + w[i].L + 1;
  parallel_for(ctx, n, win_item, &a);
  free(xoff);
  return BO_OK;
}

int bo_backend_fs_bck_decode(void *ctx, const void *wins, int n, const float xfE[2], const float xf5_loop[3], const int64_t *out_offset,
                             float *mocc, float *btot, float *etot, float *fwdsc, float *bcksc, int32_t *status)
{
  bo_backend *b = ctx;
  snprintf(b->err, sizeof b->err, "fs_bck_decode is not used by the host pipeline; not provided by the CPU backend");
  return BO_EINVAL;
}

/* ---- envelope stage ---- */
typedef struct { const bathgpu_envelope *e; const float *xfE5; bathgpu_domain_result *res; BO_TRACE **tr; } env_args;

static void env_item(bo_backend *b, void *arg, int i)
{
  env_args *a = arg;
  const bathgpu_envelope *e = &a->e[i];
  bathgpu_domain_result *r = &a->res[i];
  BO_FS_OPROFILE om = *b->om5;
  int L = e->L, st, M = om.M;
  uint8_t *sub = malloc((size_t) L + 2);
  BO_MX *fwd = bo_mx_create(M, L, 8), *bck = bo_mx_create(M, L, 3), *oa = bo_mx_create(M, L, 3);
  sub[0] = sub[L + 1] = BO_DSQ_SENTINEL;
  memcpy(sub + 1, b->dsq[b->cur] + e->start, (size_t) L);
  om.nj = 0.0f;
  om.xf[BO_X_N][BO_O_LOOP] = om.xf[BO_X_C][BO_O_LOOP] = om.xf[BO_X_J][BO_O_LOOP] = e->ploop;
  om.xf[BO_X_N][BO_O_MOVE] = om.xf[BO_X_C][BO_O_MOVE] = om.xf[BO_X_J][BO_O_MOVE] = e->pmove;
  om.xf[BO_X_E][BO_O_MOVE] = a->xfE5[0]; om.xf[BO_X_E][BO_O_LOOP] = a->xfE5[1];
  memset(r, 0, sizeof *r);
  a->tr[i] = NULL;
  st = bo_Forward_Frameshift(sub, L, &om, fwd, &r->envsc);
  if (st == BO_OK) st = bo_Backward_Frameshift(sub, L, &om, fwd, bck, &r->bcksc);
  if (st == BO_OK) st = bo_Decoding_Frameshift(&om, fwd, bck);
  if (st == BO_OK) {
    BO_TRACE *tr = bo_trace_create();
    bo_OptimalAccuracy_Frameshift(&om, fwd, oa, &r->oasc);
    if (bo_OATrace_Frameshift(&om, fwd, oa, tr) == BO_OK) a->tr[i] = tr; else { bo_trace_destroy(tr); st = BO_EINVAL; }
    bo_Null2_fs_ByExpectation(&om, fwd, r->null2);
  }
  r->status = st;
  bo_mx_destroy(fwd); bo_mx_destroy(bck); bo_mx_destroy(oa); free(sub);
}

int bo_backend_fs_domains(void *ctx, const void *envs, int n, const float xfE5[2], void *results, void *traces, int64_t max_steps)
{
  bo_backend *b = ctx;
  BO_TRACE **tr = calloc((size_t) n, sizeof(BO_TRACE *));
  bathgpu_domain_result *res = results;
  bathgpu_trace_step *out = traces;
  env_args a = { envs, xfE5, res, tr };
  int64_t used = 0;
  int i, z, rc = BO_OK;
  parallel_for(b, n, env_item, &a);
  for (i = 0; i < n; i++) {
    res[i].trace_offset = (int32_t) used; res[i].trace_len = 0;
    if (!tr[i]) continue;
    if (used + tr[i]->N > max_steps) { snprintf(b->err, sizeof b->err, "trace buffer too small"); rc = BO_EINVAL; }
    else {
      for (z = 0; z < tr[i]->N; z++) {
        out[used + z].i = tr[i]->i[z]; out[used + z].k = (int16_t) tr[i]->k[z]; out[used + z].st = (uint8_t) tr[i]->st[z];
        out[used + z].c = (uint8_t) tr[i]->c[z]; out[used + z].pp = tr[i]->pp[z];
      }
      res[i].trace_len = tr[i]->N;
      used += tr[i]->N;
    }
    bo_trace_destroy(tr[i]);
  }
  free(tr);
  return rc;
}

/* ---- Forward matrices of multi-domain regions, for the host's stochastic traceback ---- */
typedef struct { const bathgpu_envelope *e; const float *xfE5; float *mx, *xr, *sc; int32_t *st; const int64_t *off; } fmx_args;

static void fmx_item(bo_backend *b, void *arg, int i)
{
  fmx_args *a = arg;
  const bathgpu_envelope *e = &a->e[i];
  BO_FS_OPROFILE om = *b->om5;
  int L = e->L, M = om.M;
  uint8_t *sub = malloc((size_t) L + 2);
  BO_MX *fwd = bo_mx_create(M, L, 8);
  sub[0] = sub[L + 1] = BO_DSQ_SENTINEL;
  memcpy(sub + 1, b->dsq[b->cur] + e->start, (size_t) L);
  om.xf[BO_X_N][BO_O_LOOP] = om.xf[BO_X_C][BO_O_LOOP] = om.xf[BO_X_J][BO_O_LOOP] = e->ploop;
  om.xf[BO_X_N][BO_O_MOVE] = om.xf[BO_X_C][BO_O_MOVE] = om.xf[BO_X_J][BO_O_MOVE] = e->pmove;
  om.xf[BO_X_E][BO_O_MOVE] = a->xfE5[0]; om.xf[BO_X_E][BO_O_LOOP] = a->xfE5[1];
  a->sc[i] = 0.0f;
  a->st[i] = bo_Forward_Frameshift(sub, L, &om, fwd, &a->sc[i]);
  if (a->mx) memcpy(a->mx + (size_t) a->off[i] * (M + 1) * 8, fwd->dp, sizeof(float) * (size_t)(L + 1) * (M + 1) * 8);
  if (a->xr) memcpy(a->xr + (size_t) a->off[i] * 6, fwd->xmx, sizeof(float) * 6 * (size_t)(L + 1));
  bo_mx_destroy(fwd); free(sub);
}

int bo_backend_fs_forward_matrices(void *ctx, const void *regs, int n, const float xfE5[2], float *mx, float *xrows, int64_t max_rows,
                                   float *fwdsc, int32_t *status)
{
  bo_backend *b = ctx;
  const bathgpu_envelope *e = regs;
  int64_t *off = malloc(sizeof(int64_t) * (size_t)(n + 1));
  fmx_args a = { e, xfE5, mx, xrows, fwdsc, status, off };
  int i;
  off[0] = 0;
  for (i = 0; i < n; i++) off[i + 1] = off[i] + e[i].L + 1;
  if ((mx || xrows) && off[n] > max_rows) { free(off); snprintf(b->err, sizeof b->err, "matrix buffer too small"); return BO_EINVAL; }
  parallel_for(b, n, fmx_item, &a);
  free(off);
  return BO_OK;
}

/* ---- standard-translation branch (orf_domain.c) ---- */
typedef struct { const bathgpu_orf *orfs; float nj; const float *xfE; float *fsc, *bsc; int32_t *st; float *fx, *bx; const int64_t *xoff; } orfx_args;

static void set_length_model(BO_OPROFILE *om, float pmove, float ploop, const float *xfE)
{
  om->xf[BO_X_N][BO_O_LOOP] = om->xf[BO_X_C][BO_O_LOOP] = om->xf[BO_X_J][BO_O_LOOP] = ploop;
  om->xf[BO_X_N][BO_O_MOVE] = om->xf[BO_X_C][BO_O_MOVE] = om->xf[BO_X_J][BO_O_MOVE] = pmove;
  om->xf[BO_X_E][BO_O_MOVE] = xfE[0]; om->xf[BO_X_E][BO_O_LOOP] = xfE[1];
}

static void orfx_item(bo_backend *b, void *arg, int i)
{
  orfx_args *a = arg;
  const bathgpu_orf *o = &a->orfs[i];
  BO_OPROFILE om = *b->om;
  uint8_t *d = orf_dsq(b, o);
  int L = o->L, st;
  float pmove = (2.0f + a->nj) / ((float) L + 2.0f + a->nj), fsc = 0.0f, bsc = 0.0f;
  BO_MX *oxf = bo_mx_create(om.M, L, 0), *oxb = bo_mx_create(om.M, L, 0);
  set_length_model(&om, pmove, 1.0f - pmove, a->xfE);
  st = bo_Forward(d, L, &om, oxf, &fsc);
  if (st == BO_OK) st = bo_Backward(d, L, &om, oxf, oxb, &bsc);
  memcpy(a->fx + a->xoff[i] * 6, oxf->xmx, sizeof(float) * 6 * (size_t)(L + 1));
  memcpy(a->bx + a->xoff[i] * 6, oxb->xmx, sizeof(float) * 6 * (size_t)(L + 1));
  if (a->fsc) a->fsc[i] = fsc;
  if (a->bsc) a->bsc[i] = bsc;
  a->st[i] = st;
  bo_mx_destroy(oxf); bo_mx_destroy(oxb); free(d);
}

int bo_backend_orf_fwd_bck_xrows(void *ctx, const void *orfs, int n, float nj, const float xfE[2], float *fwd_xrows, float *bck_xrows,
                                 float *fwdsc, float *bcksc, int32_t *status)
{
  const bathgpu_orf *o = orfs;
  int64_t *xoff = malloc(sizeof(int64_t) * (size_t)(n + 1));
  orfx_args a = { orfs, nj, xfE, fwdsc, bcksc, status, fwd_xrows, bck_xrows, xoff };
  int i;
  xoff[0] = 0;
  for (i = 0; i < n; i++) xoff[i + 1] = xoff[i] + o[i].L + 1;
  parallel_for(ctx, n, orfx_item, &a);
  free(xoff);
  return BO_OK;
}

static void orfenv_item(bo_backend *b, void *arg, int i)
{
  env_args *a = arg;
  const bathgpu_envelope *e = &a->e[i];
  bathgpu_domain_result *r = &a->res[i];
  BO_OPROFILE om = *b->om;
  int L = e->L, st, M = om.M;
  uint8_t *sub = malloc((size_t) L + 2);
  BO_MX *fwd = bo_mx_create(M, L, 3), *bck = bo_mx_create(M, L, 3);
  sub[0] = sub[L + 1] = BO_DSQ_SENTINEL;
  memcpy(sub + 1, b->res[b->cur] + e->start, (size_t) L);
  om.nj = 0.0f;
  set_length_model(&om, e->pmove, e->ploop, a->xfE5);
  memset(r, 0, sizeof *r);
  a->tr[i] = NULL;
  st = bo_Forward(sub, L, &om, fwd, &r->envsc);
  if (st == BO_OK) st = bo_Backward(sub, L, &om, fwd, bck, &r->bcksc);
  if (st == BO_OK) st = bo_Decoding(&om, fwd, bck, bck);          /* bck now holds the posteriors (p7_domaindef.c:1251) */
  if (st == BO_OK) {
    BO_TRACE *tr = bo_trace_create();
    bo_OptimalAccuracy(&om, bck, fwd, &r->oasc);                  /* fwd now holds the OA matrix (:1255) */
    if (bo_OATrace(&om, bck, fwd, b->lanes_u8 / 4 > 0 ? b->lanes_u8 / 4 : 4, tr) == BO_OK) a->tr[i] = tr;
    else { bo_trace_destroy(tr); st = BO_EINVAL; }
    bo_Null2_ByExpectation(&om, bck, r->null2);
  }
  r->status = st;
  bo_mx_destroy(fwd); bo_mx_destroy(bck); free(sub);
}

int bo_backend_orf_domains(void *ctx, const void *envs, int n, const float xfE[2], void *results, void *traces, int64_t max_steps)
{
  bo_backend *b = ctx;
  BO_TRACE **tr = calloc((size_t) n, sizeof(BO_TRACE *));
  bathgpu_domain_result *res = results;
  bathgpu_trace_step *out = traces;
  env_args a = { envs, xfE, res, tr };
  int64_t used = 0;
  int i, z, rc = BO_OK;
  parallel_for(b, n, orfenv_item, &a);
  for (i = 0; i < n; i++) {
    res[i].trace_offset = (int32_t) used; res[i].trace_len = 0;
    if (!tr[i]) continue;
    if (used + tr[i]->N > max_steps) { snprintf(b->err, sizeof b->err, "trace buffer too small"); rc = BO_EINVAL; }
    else {
      for (z = 0; z < tr[i]->N; z++) {
        out[used + z].i = tr[i]->i[z]; out[used + z].k = (int16_t) tr[i]->k[z]; out[used + z].st = (uint8_t) tr[i]->st[z];
        out[used + z].c = (uint8_t) tr[i]->c[z]; out[used + z].pp = tr[i]->pp[z];
      }
      res[i].trace_len = tr[i]->N;
      used += tr[i]->N;
    }
    bo_trace_destroy(tr[i]);
  }
  free(tr);
  return rc;
}

/* ---- translation + MSV + F1 screen over all blocks of the resident strand (bathgpu_orfs_msv_screen on the CPU) ---- */
typedef struct { const bathgpu_block *blk; int complement; const uint8_t *gcode; int min_len; BO_ORF **orfs; int *norf; uint8_t **res; int64_t *nres; } xl_args;
static void xl_item(bo_backend *b, void *arg, int i)
{
  xl_args *a = arg;
  a->orfs[i] = NULL; a->norf[i] = 0; a->res[i] = NULL; a->nres[i] = 0;
  if (a->blk[i].n >= 3) bo_find_orfs(b->dsq[b->cur] + a->blk[i].goff, a->blk[i].n, a->gcode, a->min_len, &a->orfs[i], &a->norf[i], &a->res[i], &a->nres[i]);
}
/* MSV over every ORF of every block in ONE parallel pass (item = global ORF index; blk_of / first map it back to its block) */
typedef struct { BO_ORF **orfs; uint8_t **res; const int *blk_of; const int64_t *first; const uint8_t *tjb_of; int max_len; float *usc; int32_t *st; const uint8_t *dead; } msvx_args;
static void msvx_item(bo_backend *b, void *arg, int g)
{
  msvx_args *a = arg;
  BO_OPROFILE om = *b->om;
  const int bi = a->blk_of[g], i = (int) (g - a->first[bi]);
  const BO_ORF *o = &a->orfs[bi][i];
  int L = o->n;
  uint8_t stackbuf[2048], *d;
  if (a->dead[g]) { a->usc[g] = -INFINITY; a->st[g] = 0; return; }
  d = (L + 2 <= (int) sizeof stackbuf) ? stackbuf : malloc((size_t) L + 2);
  d[0] = d[L + 1] = BO_DSQ_SENTINEL;
  memcpy(d + 1, a->res[bi] + o->offset, (size_t) L);
  om.tjb_b = a->tjb_of[L < a->max_len ? L : a->max_len];
  if (b->use_simd && b->msv_simd && om.M <= 1024) a->st[g] = bo_MSVFilter_simd(b->msv_simd, d, L, &om, &a->usc[g]);
  else a->st[g] = bo_MSVFilter(d, L, &om, &a->usc[g]);
  if (d != stackbuf) free(d);
}


int bo_backend_orfs_msv_screen(void *ctx, const void *blocks, int nblocks, int complement, const uint8_t gcode[64], int min_len,
                               const uint8_t *tjb_of, const float *null_of, int max_len, double min_bits,
                               int64_t *norfs_per_block, int64_t *nhits, int64_t *nres)
{
  bo_backend *b = ctx;
  const bathgpu_block *blk = blocks;
  BO_ORF **orfs = calloc((size_t) nblocks, sizeof(BO_ORF *));
  int *norf = calloc((size_t) nblocks, sizeof(int));
  uint8_t **res = calloc((size_t) nblocks, sizeof(uint8_t *));
  int64_t *nr = calloc((size_t) nblocks, sizeof(int64_t));
  int64_t *first = calloc((size_t) nblocks + 1, sizeof(int64_t));
  xl_args xa = { blk, complement, gcode, min_len, orfs, norf, res, nr };
  int64_t tot_hits = 0, tot_res = 0, cap_res = 0, cap_hits = 0, N, g;
  int bi, i, s = b->cur;
  int *blk_of; float *usc; int32_t *st; uint8_t *dead;
  parallel_for(b, nblocks, xl_item, &xa);
  free(b->hits[s]); b->hits[s] = NULL; b->nhits[s] = 0;
  free(b->res[s]); b->res[s] = NULL; b->nres[s] = 0;
  for (bi = 0; bi < nblocks; bi++) { cap_res += nr[bi]; first[bi + 1] = first[bi] + norf[bi]; if (norfs_per_block) norfs_per_block[bi] = norf[bi]; }
  N = first[nblocks];
  if (N > 0x7fffffffLL) { snprintf(b->err, sizeof b->err, "too many ORFs in one call"); return BO_EINVAL; }
  b->res[s] = malloc((size_t) (cap_res > 0 ? cap_res : 1));
  blk_of = malloc(sizeof(int) * (size_t) (N > 0 ? N : 1));
  usc = malloc(sizeof(float) * (size_t) (N > 0 ? N : 1));
  st = malloc(sizeof(int32_t) * (size_t) (N > 0 ? N : 1));
  dead = calloc((size_t) (N > 0 ? N : 1), 1);
  for (bi = 0; bi < nblocks; bi++)
    for (i = 0; i < norf[bi]; i++) {
      g = first[bi] + i;
      blk_of[g] = bi;
      dead[g] = complement ? ((blk[bi].n - orfs[bi][i].start + 1) < blk[bi].C) : (orfs[bi][i].end < blk[bi].C);
    }
  {
    msvx_args ma = { orfs, res, blk_of, first, tjb_of, max_len, usc, st, dead };
    parallel_for(b, (int) N, msvx_item, &ma);
  }
  for (bi = 0; bi < nblocks; bi++) {
    for (i = 0; i < norf[bi]; i++) {
      int L = orfs[bi][i].n, keep;
      g = first[bi] + i;
      if (dead[g]) continue;
      keep = (st[g] != 0) || (((double) usc[g] - (double) null_of[L < max_len ? L : max_len]) / 0.69314718055994529 >= min_bits);
      if (!keep) continue;
      if (tot_hits == cap_hits) { cap_hits = cap_hits ? 2 * cap_hits : 1024; b->hits[s] = realloc(b->hits[s], sizeof(bathgpu_orf_hit) * (size_t) cap_hits); }
      b->hits[s][tot_hits].block = bi; b->hits[s][tot_hits].index = i; b->hits[s][tot_hits].start = orfs[bi][i].start;
      b->hits[s][tot_hits].end = orfs[bi][i].end; b->hits[s][tot_hits].n = L; b->hits[s][tot_hits].frame = orfs[bi][i].frame;
      b->hits[s][tot_hits].offset = tot_res; b->hits[s][tot_hits].usc = usc[g]; b->hits[s][tot_hits].status = st[g];
      memcpy(b->res[s] + tot_res, res[bi] + orfs[bi][i].offset, (size_t) L);
      tot_res += L; tot_hits++;
    }
    free(orfs[bi]); free(res[bi]);
  }
  free(usc); free(st); free(dead); free(blk_of); free(first);
  free(orfs); free(norf); free(res); free(nr);
  b->nhits[s] = tot_hits; b->nres[s] = tot_res;
  *nhits = tot_hits; *nres = tot_res;
  return BO_OK;
}

int bo_backend_orfs_fetch(void *ctx, void *hits, uint8_t *residues)
{
  bo_backend *b = ctx;
  int s = b->cur;
  if (b->nhits[s] > 0) {
    memcpy(hits, b->hits[s], sizeof(bathgpu_orf_hit) * (size_t) b->nhits[s]);
    memcpy(residues, b->res[s], (size_t) b->nres[s]);
  }
  return BO_OK;
}

/* plain heap memory: the CPU backend has no device link to feed */
void *bo_backend_host_alloc(size_t bytes) { return malloc(bytes); }
void  bo_backend_host_free(void *p) { free(p); }

/* ---- a5: the bias-composition filter (bathgpu_bias_forward) ---------------------------------------------------------------
 * esl_hmm_Forward over the 2-state filter HMM that p7_bg_SetFilter configures (src/p7_bg.c:449-471), as p7_bg_FilterScore runs it
 * on an ORF (:491-500) and p7_bg_fs_FilterScore on the three reading frames of a DNA window (:522-573).  Easel is absent from
 * /root/reference (INSTALL:6-8), so esl_hmm_Forward is restated from Easel's published esl_hmm.c: row i holds
 * fwd[i][k] = e_k(x_i) * sum_m fwd[i-1][m] t[m][k] divided by the row maximum, sc[i] = log(max), the termination row adds
 * log(sum_m fwd[L][m] t[m][E]) with t[m][E] = 1, and the score is the float sum of sc[] in row order. */
static float bias_hmm_forward(const float *eo, float t00, float t10, float t11, const uint8_t *x, int L)
{
  float t[2][2], pi[2] = { 0.999, 0.001 }, prev[2], cur[2], max, logsc = 0.0f, last;
  int   i, k, m;
  if (L == 0) return 0.0f;
  t[0][0] = t00; t[0][1] = 1.0f - t00; t[1][0] = t10; t[1][1] = t11;
  max = 0.0;
  for (k = 0; k < 2; k++) { prev[k] = eo[2 * x[0] + k] * pi[k]; if (prev[k] > max) max = prev[k]; }
  for (k = 0; k < 2; k++) prev[k] /= max;
  logsc += (float) log(max);
  for (i = 1; i < L; i++) {
    max = 0.0;
    for (k = 0; k < 2; k++) {
      cur[k] = 0.0;
      for (m = 0; m < 2; m++) cur[k] += prev[m] * t[m][k];
      cur[k] *= eo[2 * x[i] + k];
      if (cur[k] > max) max = cur[k];
    }
    for (k = 0; k < 2; k++) prev[k] = cur[k] / max;
    logsc += (float) log(max);
  }
  last = 0.0;
  for (m = 0; m < 2; m++) last += prev[m] * 1.0f;
  logsc += (float) log(last);
  return logsc;
}

typedef struct { int kind; const bathgpu_bias_item *items; const float *tables; float t10, t11; const uint8_t *gcode; float *out; } bias_job;
static void bias_item_fn(bo_backend *b, void *arg, int it)
{
  bias_job *j = arg;
  const bathgpu_bias_item *d = &j->items[it];
  const float *eo = j->tables + (size_t) d->table * 58;
  if (j->kind == 0) j->out[it] = bias_hmm_forward(eo, d->t00, j->t10, j->t11, b->res[b->cur] + d->start, d->L);
  else {
    const uint8_t *dna = b->dsq[b->cur] + d->start - 1;          /* window position p is dna[p] */
    uint8_t *orf = malloc((size_t) d->L + 2);
    int f, i, n;
    for (f = 1; f <= 3; f++) {
      n = 0;
      for (i = f; i <= d->L - 2; i += 3) {
        uint8_t a = dna[i], c = dna[i + 1], g = dna[i + 2], aa;
        if (a >= 4 || c >= 4 || g >= 4) continue;               /* a codon with a degenerate nucleotide is X here, as in the ORF finder */
        aa = j->gcode[16 * a + 4 * c + g];
        if (aa < 20) orf[n++] = aa;
      }
      j->out[3 * it + f - 1] = bias_hmm_forward(eo, d->t00, j->t10, j->t11, orf, n);
    }
    free(orf);
  }
}
int bo_backend_bias_forward(void *ctx, int kind, const void *items, int n, const float *tables, int ntab, float t10, float t11,
                            const uint8_t gcode[64], float *out)
{
  bo_backend *b = ctx;
  bias_job j;
  if (!b || (kind != 0 && kind != 1) || n < 0 || ntab < 1) return BO_EINVAL;
  if ((kind == 0 && !b->res[b->cur]) || (kind == 1 && !b->dsq[b->cur])) return BO_EINVAL;
  j.kind = kind; j.items = items; j.tables = tables; j.t10 = t10; j.t11 = t11; j.gcode = gcode; j.out = out;
  parallel_for(b, n, bias_item_fn, &j);
  return BO_OK;
}

/* ---- f2, multi-domain regions of the standard branch (bathgpu_orf_forward_matrices): p7_Forward over a region of an ORF in multihit
 * mode at the ORF's length model, whole matrix handed back (src/p7_domaindef.c:561-562): cells {M, D, I, 0} per (row, node) */
typedef struct { const bathgpu_envelope *e; const float *xfE; float *mx, *xr, *sc; int32_t *st; const int64_t *off; } orffm_args;
static void orffm_item(bo_backend *b, void *arg, int i)
{
  orffm_args *a = arg;
  const bathgpu_envelope *e = &a->e[i];
  BO_OPROFILE om = *b->om;
  int L = e->L, M = om.M, r, k;
  uint8_t *sub = malloc((size_t) L + 2);
  BO_MX *fwd = bo_mx_create(M, L, 3);
  float *mx = a->mx + (size_t) a->off[i] * (M + 1) * 4, *xr = a->xr + (size_t) a->off[i] * 6;
  sub[0] = sub[L + 1] = BO_DSQ_SENTINEL;
  memcpy(sub + 1, b->res[b->cur] + e->start, (size_t) L);
  om.nj = 1.0f;
  set_length_model(&om, e->pmove, e->ploop, a->xfE);
  a->st[i] = bo_Forward(sub, L, &om, fwd, &a->sc[i]);
  for (r = 0; r <= L; r++) {
    for (k = 0; k <= M; k++) {
      float *c = mx + ((size_t) r * (M + 1) + k) * 4;
      c[0] = fwd->dp[((size_t) r * (M + 1) + k) * 3 + BO_S_M]; c[1] = fwd->dp[((size_t) r * (M + 1) + k) * 3 + BO_S_D];
      c[2] = fwd->dp[((size_t) r * (M + 1) + k) * 3 + BO_S_I]; c[3] = 0.0f;
    }
    memcpy(xr + (size_t) r * 6, fwd->xmx + (size_t) r * 6, sizeof(float) * 6);
  }
  bo_mx_destroy(fwd); free(sub);
}
int bo_backend_orf_forward_matrices(void *ctx, const void *regs, int n, const float xfE[2], float *mx, float *xrows, int64_t max_rows,
                                    float *fwdsc, int32_t *status)
{
  bo_backend *b = ctx;
  const bathgpu_envelope *e = regs;
  int64_t *off = malloc(sizeof(int64_t) * (size_t) (n + 1));
  orffm_args a = { e, xfE, mx, xrows, fwdsc, status, off };
  int i;
  off[0] = 0;
  for (i = 0; i < n; i++) off[i + 1] = off[i] + e[i].L + 1;
  if (off[n] > max_rows) { free(off); snprintf(b->err, sizeof b->err, "matrix buffer too small"); return BO_EINVAL; }
  parallel_for(b, n, orffm_item, &a);
  free(off);
  return BO_OK;
}
