/* orfs.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 * Six-frame translation as bathsearch asks Easel for it (esl_gencode_ProcessStart/Piece/End, Easel branch BATH, not in the
 * tree; call site src/bathsearch.c:385-392 with the work state of :819): on one strand, three frames; an ORF is a maximal
 * stop-free run of whole codons with at least min_len residues; any codon may start one; a codon holding a degenerate
 * nucleotide translates to X.  ORFs come out in order of their last nucleotide, as a left-to-right scan finishes them.
 * Pinned only through the golden outputs (footer counters and hits of tutorial/ *.out). */
#include <stdlib.h>
#include <string.h>
#include "bath_oracle.h"

static int by_end(const void *a, const void *b) { return ((const BO_ORF *) a)->end - ((const BO_ORF *) b)->end; }

/* dsq[1..n]; gcode[64] amino codes with BO_AA_STOP for stops.  Returns ORFs (malloc'd) and their residues, concatenated. */
int bo_find_orfs(const uint8_t *dsq, int n, const uint8_t *gcode, int min_len, BO_ORF **ret_orfs, int *ret_n, uint8_t **ret_res, int64_t *ret_nres)
{
  int      cap = 64, norf = 0, f, i;
  int64_t  nres = 0;
  BO_ORF  *orfs = malloc(sizeof(BO_ORF) * cap);
  uint8_t *res = malloc((size_t) (n > 0 ? n : 1));
  if (!orfs || !res) { free(orfs); free(res); return BO_EMEM; }
  for (f = 0; f < 3; f++) {
    int     run_start = -1;
    int64_t run_res = nres;
    for (i = f + 1; ; i += 3) {
      int stop_here = 0, at_end = (i + 2 > n);
      if (!at_end) {
        uint8_t a = dsq[i], b = dsq[i+1], c = dsq[i+2], aa = BO_AA_X;
        if (a < 4 && b < 4 && c < 4) aa = gcode[16 * a + 4 * b + c];
        if (aa == BO_AA_STOP) stop_here = 1;
        else { if (run_start < 0) { run_start = i; run_res = nres; } res[nres++] = aa; }
      }
      if (stop_here || at_end) {
        int len = (int) (nres - run_res);
        if (run_start > 0 && len >= min_len) {
          if (norf == cap) { cap *= 2; orfs = realloc(orfs, sizeof(BO_ORF) * cap); if (!orfs) { free(res); return BO_EMEM; } }
          orfs[norf].start = run_start; orfs[norf].end = run_start + 3 * len - 1; orfs[norf].frame = f;
          orfs[norf].n = len; orfs[norf].offset = run_res;
          norf++;
        } else nres = run_res;
        run_start = -1; run_res = nres;
      }
      if (at_end) break;
    }
  }
  qsort(orfs, (size_t) norf, sizeof(BO_ORF), by_end);
  *ret_orfs = orfs; *ret_n = norf; *ret_res = res; *ret_nres = nres;
  return BO_OK;
}
