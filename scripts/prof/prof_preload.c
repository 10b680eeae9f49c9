/* scratch tuning aid: a sampling profiler for the host pipeline when no perf / gdb is around.
 *   gcc -O2 -shared -fPIC -o /tmp/libbprof.so scripts/prof/prof_preload.c
 *   BPROF=/tmp/bprof.out LD_PRELOAD=/tmp/libbprof.so python <script>;  python scripts/prof/prof_report.py /tmp/bprof.out
 * ITIMER_PROF ticks every millisecond of process CPU time; the handler records the interrupted instruction pointer (whatever thread
 * was running), and the destructor writes the samples with /proc/self/maps so that the report can symbolise them with nm. */
#define _GNU_SOURCE
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <ucontext.h>

#define CAP (1 << 21)
static void *samples[CAP];
static volatile int nsamples;

static void on_tick(int sig, siginfo_t *si, void *uc_)
{
  ucontext_t *uc = uc_;
  int i = __sync_fetch_and_add(&nsamples, 1);
  (void) sig; (void) si;
  if (i < CAP) samples[i] = (void *) uc->uc_mcontext.gregs[REG_RIP];
}

__attribute__((constructor)) static void bprof_init(void)
{
  struct sigaction sa;
  struct itimerval it = { { 0, 1000 }, { 0, 1000 } };
  if (!getenv("BPROF")) return;
  memset(&sa, 0, sizeof sa);
  sa.sa_sigaction = on_tick; sa.sa_flags = SA_SIGINFO | SA_RESTART;
  sigaction(SIGPROF, &sa, NULL);
  setitimer(ITIMER_PROF, &it, NULL);
}

__attribute__((destructor)) static void bprof_fini(void)
{
  const char *path = getenv("BPROF");
  struct itimerval off = { { 0, 0 }, { 0, 0 } };
  FILE *f, *m;
  char line[1024];
  int i, n;
  if (!path) return;
  setitimer(ITIMER_PROF, &off, NULL);
  n = nsamples < CAP ? nsamples : CAP;
  if (n == 0) return;                       /* a wrapper process (timeout, sh) that did no work */
  f = fopen(path, "w");
  if (!f) return;
  m = fopen("/proc/self/maps", "r");
  if (m) { while (fgets(line, sizeof line, m)) if (strstr(line, " r-xp ")) fprintf(f, "M %s", line); fclose(m); }
  for (i = 0; i < n; i++) fprintf(f, "S %p\n", samples[i]);
  fclose(f);
}
