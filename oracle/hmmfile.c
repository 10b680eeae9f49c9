/* hmmfile.c -- ORACLE (test infrastructure only).
 * Minimal reader for BATH3/f (== HMMER3/f) ASCII profile files.
 * Follows src/p7_hmmfile.c:1374-1690 (read_asc30hmm): header tags NAME, ACC, LENG,
 * MAXL, STATS LOCAL {MSV,VITERBI,FORWARD,FS3 FORWARD,FS5 FORWARD}, FRAMESHIFT PROB,
 * CODON TABLE; body values are -ln p, '*' = 0, converted with expf(-1.0*atof(tok)).
 * Quirk kept (p7_hmmfile.c:1509-1510): on "STATS LOCAL FS3 FORWARD tau lambda"
 * only tau is read; FLAMBDA comes from the FORWARD line. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "bath_oracle.h"

#define LINEMAX 16384

static float prob_tok(const char *tok)
{
  return (*tok == '*') ? 0.0f : expf(-1.0 * atof(tok));
}

void bo_hmm_destroy(BO_HMM *hmm)
{
  if (!hmm) return;
  free(hmm->t); free(hmm->mat); free(hmm->ins); free(hmm->consensus); free(hmm);
}

static int read_one(FILE *fp, BO_HMM **ret_hmm)
{
  char  line[LINEMAX];
  char *tok, *save;
  BO_HMM *hmm = NULL;
  int   k, x, z;
  int   seen_magic = 0;

  /* find magic line */
  while (fgets(line, LINEMAX, fp)) {
    tok = strtok_r(line, " \t\r\n", &save);
    if (!tok) continue;
    if (strncmp(tok, "BATH3/", 6) == 0 || strncmp(tok, "HMMER3/", 7) == 0) { seen_magic = 1; break; }
    return BO_EFORMAT;
  }
  if (!seen_magic) return BO_EOF;

  hmm = calloc(1, sizeof(BO_HMM));
  if (!hmm) return BO_EMEM;
  for (z = 0; z < 8; z++) hmm->evparam[z] = -99999.0f;  /* p7_EVPARAM_UNSET */
  hmm->fsprob = -1.0f; hmm->ct = -1; hmm->max_length = -1;

  while (fgets(line, LINEMAX, fp)) {
    tok = strtok_r(line, " \t\r\n", &save);
    if (!tok) continue;
    if      (strcmp(tok, "NAME") == 0) { tok = strtok_r(NULL, " \t\r\n", &save); if (tok) strncpy(hmm->name, tok, 127); }
    else if (strcmp(tok, "ACC")  == 0) { tok = strtok_r(NULL, " \t\r\n", &save); if (tok) strncpy(hmm->acc, tok, 63); }
    else if (strcmp(tok, "LENG") == 0) { tok = strtok_r(NULL, " \t\r\n", &save); hmm->M = atoi(tok); }
    else if (strcmp(tok, "MAXL") == 0) { tok = strtok_r(NULL, " \t\r\n", &save); hmm->max_length = atoi(tok); }
    else if (strcmp(tok, "ALPH") == 0) { tok = strtok_r(NULL, " \t\r\n", &save);
                                         if (!tok || strcmp(tok, "amino") != 0) { bo_hmm_destroy(hmm); return BO_EFORMAT; } }
    else if (strcmp(tok, "STATS") == 0) {
      char *t1 = strtok_r(NULL, " \t\r\n", &save);   /* LOCAL */
      char *t2 = strtok_r(NULL, " \t\r\n", &save);   /* MSV | VITERBI | FORWARD | FS3 | FS5 */
      char *t3 = strtok_r(NULL, " \t\r\n", &save);
      char *t4 = strtok_r(NULL, " \t\r\n", &save);
      if (!t1 || !t2 || !t3 || !t4 || strcmp(t1, "LOCAL") != 0) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
      if      (strcmp(t2, "MSV")     == 0) { hmm->evparam[BO_MMU]  = atof(t3); hmm->evparam[BO_MLAMBDA] = atof(t4); }
      else if (strcmp(t2, "VITERBI") == 0) { hmm->evparam[BO_VMU]  = atof(t3); hmm->evparam[BO_VLAMBDA] = atof(t4); }
      else if (strcmp(t2, "FORWARD") == 0) { hmm->evparam[BO_FTAU] = atof(t3); hmm->evparam[BO_FLAMBDA] = atof(t4); }
      else if (strcmp(t2, "FS3")     == 0) { hmm->evparam[BO_FTAUFS3] = atof(t4); hmm->has_stats_fs3 = 1; }
      else if (strcmp(t2, "FS5")     == 0) { hmm->evparam[BO_FTAUFS5] = atof(t4); hmm->has_stats_fs5 = 1; }
    }
    else if (strcmp(tok, "FRAMESHIFT") == 0) { strtok_r(NULL, " \t\r\n", &save); tok = strtok_r(NULL, " \t\r\n", &save); if (tok) hmm->fsprob = atof(tok); }
    else if (strcmp(tok, "CODON") == 0)      { strtok_r(NULL, " \t\r\n", &save); tok = strtok_r(NULL, " \t\r\n", &save); if (tok) hmm->ct = atoi(tok); }
    else if (strcmp(tok, "HMM") == 0) break;
  }
  if (hmm->M <= 0) { bo_hmm_destroy(hmm); return BO_EFORMAT; }

  /* skip the transition header line */
  if (!fgets(line, LINEMAX, fp)) { bo_hmm_destroy(hmm); return BO_EFORMAT; }

  hmm->t   = calloc((size_t)(hmm->M + 1) * 7,    sizeof(float));
  hmm->mat = calloc((size_t)(hmm->M + 1) * BO_K, sizeof(float));
  hmm->ins = calloc((size_t)(hmm->M + 1) * BO_K, sizeof(float));
  hmm->consensus = calloc(hmm->M + 2, 1);
  if (!hmm->t || !hmm->mat || !hmm->ins || !hmm->consensus) { bo_hmm_destroy(hmm); return BO_EMEM; }
  hmm->consensus[0] = ' ';

  if (!fgets(line, LINEMAX, fp)) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
  tok = strtok_r(line, " \t\r\n", &save);
  if (tok && strcmp(tok, "COMPO") == 0) {
    for (x = 0; x < BO_K; x++) { tok = strtok_r(NULL, " \t\r\n", &save); hmm->compo[x] = prob_tok(tok); }
    hmm->has_compo = 1;
    if (!fgets(line, LINEMAX, fp)) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
    tok = strtok_r(line, " \t\r\n", &save);
  }
  /* node 0: insert emissions, then transitions */
  for (x = 0; x < BO_K; x++) {
    if (!tok) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
    hmm->ins[x] = prob_tok(tok);
    tok = strtok_r(NULL, " \t\r\n", &save);
  }
  if (!fgets(line, LINEMAX, fp)) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
  tok = strtok_r(line, " \t\r\n", &save);
  for (x = 0; x < 7; x++) {
    if (!tok) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
    hmm->t[x] = prob_tok(tok);
    tok = strtok_r(NULL, " \t\r\n", &save);
  }

  for (k = 1; k <= hmm->M; k++) {
    if (!fgets(line, LINEMAX, fp)) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
    tok = strtok_r(line, " \t\r\n", &save);
    if (!tok || atoi(tok) != k) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
    for (x = 0; x < BO_K; x++) {
      tok = strtok_r(NULL, " \t\r\n", &save);
      if (!tok) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
      hmm->mat[k * BO_K + x] = prob_tok(tok);
    }
    tok = strtok_r(NULL, " \t\r\n", &save);           /* MAP  */
    tok = strtok_r(NULL, " \t\r\n", &save);           /* CONS */
    hmm->consensus[k] = tok ? *tok : '-';
    if (!fgets(line, LINEMAX, fp)) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
    tok = strtok_r(line, " \t\r\n", &save);
    for (x = 0; x < BO_K; x++) {
      if (!tok) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
      hmm->ins[k * BO_K + x] = prob_tok(tok);
      tok = strtok_r(NULL, " \t\r\n", &save);
    }
    if (!fgets(line, LINEMAX, fp)) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
    tok = strtok_r(line, " \t\r\n", &save);
    for (x = 0; x < 7; x++) {
      if (!tok) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
      hmm->t[k * 7 + x] = prob_tok(tok);
      tok = strtok_r(NULL, " \t\r\n", &save);
    }
  }
  /* closing // */
  if (!fgets(line, LINEMAX, fp)) { bo_hmm_destroy(hmm); return BO_EFORMAT; }
  tok = strtok_r(line, " \t\r\n", &save);
  if (!tok || strcmp(tok, "//") != 0) { bo_hmm_destroy(hmm); return BO_EFORMAT; }

  *ret_hmm = hmm;
  return BO_OK;
}

int bo_hmmfile_read(const char *path, int index, BO_HMM **ret_hmm)
{
  FILE *fp = fopen(path, "r");
  int   n, status = BO_EOF;
  BO_HMM *hmm = NULL;
  if (!fp) return BO_FAIL;
  for (n = 0; n <= index; n++) {
    hmm = NULL;
    status = read_one(fp, &hmm);
    if (status != BO_OK) break;
    if (n < index) bo_hmm_destroy(hmm);
  }
  fclose(fp);
  if (status == BO_OK) *ret_hmm = hmm;
  return status;
}

int bo_hmmfile_count(const char *path)
{
  FILE *fp = fopen(path, "r");
  int   n = 0;
  BO_HMM *hmm;
  if (!fp) return -1;
  while (read_one(fp, &hmm) == BO_OK) { bo_hmm_destroy(hmm); n++; }
  fclose(fp);
  return n;
}
