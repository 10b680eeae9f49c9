/* mx.c -- ORACLE (test infrastructure only).  DP matrix and trace containers.
 * Layout is un-striped: dp[(i*(M+1) + k)*nscells + s], xmx[i*6 + s]
 * (the reference's P7_OMX is striped for SSE lanes, impl_sse.h:329-358; striping
 * carries no semantics).  Trace follows p7_trace.c's fs variants: st,k,i,c,pp per step. */
#include <stdlib.h>
#include <string.h>
#include "bath_oracle.h"

BO_MX *bo_mx_create(int M, int L, int nscells)
{
  BO_MX *mx = calloc(1, sizeof(BO_MX));
  if (!mx) return NULL;
  mx->M = M; mx->L = L; mx->allocL = L; mx->nscells = nscells;
  if (nscells > 0) {
    mx->dp = calloc((size_t)(L + 1) * (M + 1) * nscells, sizeof(float));
    if (!mx->dp) { free(mx); return NULL; }
  }
  mx->xmx = calloc((size_t)(L + 2) * BO_NXCELLS, sizeof(float));
  if (!mx->xmx) { free(mx->dp); free(mx); return NULL; }
  return mx;
}

void bo_mx_destroy(BO_MX *mx)
{
  if (!mx) return;
  free(mx->dp); free(mx->xmx); free(mx);
}

BO_TRACE *bo_trace_create(void)
{
  BO_TRACE *tr = calloc(1, sizeof(BO_TRACE));
  if (!tr) return NULL;
  tr->nalloc = 256;
  tr->st = malloc(tr->nalloc);
  tr->k  = malloc(sizeof(int) * tr->nalloc);
  tr->i  = malloc(sizeof(int) * tr->nalloc);
  tr->c  = malloc(sizeof(int) * tr->nalloc);
  tr->pp = malloc(sizeof(float) * tr->nalloc);
  return tr;
}

void bo_trace_reuse(BO_TRACE *tr) { tr->N = 0; tr->M = 0; tr->L = 0; }

void bo_trace_destroy(BO_TRACE *tr)
{
  if (!tr) return;
  free(tr->st); free(tr->k); free(tr->i); free(tr->c); free(tr->pp); free(tr);
}

/* p7_trace_fs_AppendWithPP (src/p7_trace.c): N/C/J emit-on-transition; k is 0
 * for non-MDI states; i is 0 for non-emitting steps; c only for M. */
int bo_trace_append(BO_TRACE *tr, char st, int k, int i, int c, float pp)
{
  if (tr->N == tr->nalloc) {
    tr->nalloc *= 2;
    tr->st = realloc(tr->st, tr->nalloc);
    tr->k  = realloc(tr->k,  sizeof(int) * tr->nalloc);
    tr->i  = realloc(tr->i,  sizeof(int) * tr->nalloc);
    tr->c  = realloc(tr->c,  sizeof(int) * tr->nalloc);
    tr->pp = realloc(tr->pp, sizeof(float) * tr->nalloc);
    if (!tr->st || !tr->k || !tr->i || !tr->c || !tr->pp) return BO_EMEM;
  }
  switch (st) {
  case BO_ST_N: case BO_ST_C: case BO_ST_J:
    tr->i[tr->N]  = ((tr->N > 0 && tr->st[tr->N - 1] == st) ? i : 0);
    tr->pp[tr->N] = ((tr->N > 0 && tr->st[tr->N - 1] == st) ? pp : 0.0f);
    tr->k[tr->N]  = 0; tr->c[tr->N] = 0;
    break;
  case BO_ST_X: case BO_ST_S: case BO_ST_B: case BO_ST_E: case BO_ST_T:
    tr->i[tr->N] = 0; tr->pp[tr->N] = 0.0f; tr->k[tr->N] = 0; tr->c[tr->N] = 0;
    break;
  case BO_ST_D:
    tr->i[tr->N] = 0; tr->pp[tr->N] = 0.0f; tr->k[tr->N] = k; tr->c[tr->N] = 0;
    break;
  case BO_ST_M:
    tr->i[tr->N] = i; tr->pp[tr->N] = pp; tr->k[tr->N] = k; tr->c[tr->N] = c;
    break;
  case BO_ST_I:
    tr->i[tr->N] = i; tr->pp[tr->N] = pp; tr->k[tr->N] = k; tr->c[tr->N] = 0;
    break;
  default: return BO_EINVAL;
  }
  tr->st[tr->N] = st;
  tr->N++;
  return BO_OK;
}

/* p7_trace_fs_Reverse (src/p7_trace.c:2527-2568): pull N/C/J residues back by
 * one, then reverse in place. */
void bo_trace_reverse(BO_TRACE *tr)
{
  int z;
  for (z = 0; z + 1 < tr->N; z++) {
    if ((tr->st[z] == BO_ST_N && tr->st[z+1] == BO_ST_N) ||
        (tr->st[z] == BO_ST_C && tr->st[z+1] == BO_ST_C) ||
        (tr->st[z] == BO_ST_J && tr->st[z+1] == BO_ST_J)) {
      if (tr->i[z] == 0 && tr->i[z+1] > 0) {
        tr->i[z]  = tr->i[z+1];  tr->i[z+1]  = 0;
        tr->pp[z] = tr->pp[z+1]; tr->pp[z+1] = 0.0f;
      }
    }
  }
  for (z = 0; z < tr->N / 2; z++) {
    int   y = tr->N - z - 1;
    char  ts = tr->st[y]; tr->st[y] = tr->st[z]; tr->st[z] = ts;
    int   t;
    float tf;
    t = tr->k[y]; tr->k[y] = tr->k[z]; tr->k[z] = t;
    t = tr->i[y]; tr->i[y] = tr->i[z]; tr->i[z] = t;
    t = tr->c[y]; tr->c[y] = tr->c[z]; tr->c[z] = t;
    tf = tr->pp[y]; tr->pp[y] = tr->pp[z]; tr->pp[z] = tf;
  }
}
