/* batch.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 *
 * Runs the oracle's per-window functions over a list of windows with a pool of
 * POSIX threads, the way the reference's worker threads each own a private copy of
 * the profile and matrices (src/bathsearch.c:814-844, pipeline_thread :1224).
 * Used by tests (to check many windows quickly) and by bench.py's cpu_baseline /
 * --impl reference legs (the CPU arm timed on the box's host cores).
 */
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <xmmintrin.h>
#include "bath_oracle.h"

typedef struct {
  const uint8_t        *dsq;      /* whole block, 1-based with sentinels */
  const int64_t        *start;    /* [n] 1-based window starts */
  const int32_t        *L;        /* [n] */
  int                   n;
  const BO_FS_OPROFILE *om;
  float                *sc;       /* [n] out */
  int32_t              *status;   /* [n] out */
  int                   next;     /* work counter (atomic) */
  int                   maxL;
} fwd_job;

static void *fwd_worker(void *arg)
{
  fwd_job *job = (fwd_job *) arg;
  BO_FS_OPROFILE om = *job->om;                 /* private length model; tables shared read-only */
  BO_MX *ox = bo_mx_create(om.M, job->maxL, 0);
  uint8_t *sub = malloc((size_t) job->maxL + 2);
  _mm_setcsr(_mm_getcsr() | 0x8040);            /* flush-to-zero + denormals-are-zero in every worker, as impl_Init does (src/impl_sse/impl_sse.h:559-577; src/bathsearch.c:1235) */
  for (;;) {
    const int w = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);      /* lock-free work counter */
    if (w >= job->n) break;
    int L = job->L[w];
    /* the reference hands each window to the kernel as its own sub-sequence (src/p7_pipeline.c:1376-1380) */
    sub[0] = BO_DSQ_SENTINEL;
    memcpy(sub + 1, job->dsq + job->start[w], (size_t) L);
    sub[L + 1] = BO_DSQ_SENTINEL;
    bo_fs_oprofile_ReconfigLength(&om, L / 3);  /* src/p7_pipeline.c:1449 */
    job->status[w] = bo_ForwardParser_Frameshift_3Codons(sub, L, &om, ox, &job->sc[w]);
  }
  free(sub);
  bo_mx_destroy(ox);
  return NULL;
}

/* Forward parser (3 codon lengths) over n windows of one block with nthreads workers. */
int bo_batch_ForwardParser_3Codons(const uint8_t *dsq, const int64_t *start, const int32_t *L, int n,
                                   const BO_FS_OPROFILE *om, int nthreads, float *sc, int32_t *status)
{
  fwd_job job;
  pthread_t *th;
  int t, maxL = 0;
  if (n < 1 || nthreads < 1) return BO_EINVAL;
  for (t = 0; t < n; t++) if (L[t] > maxL) maxL = L[t];
  job.dsq = dsq; job.start = start; job.L = L; job.n = n; job.om = om; job.sc = sc; job.status = status;
  job.next = 0; job.maxL = maxL;
  th = malloc(sizeof(pthread_t) * (size_t) nthreads);
  for (t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, fwd_worker, &job);
  for (t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  return BO_OK;
}
