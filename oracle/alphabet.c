/* alphabet.c -- ORACLE (test infrastructure only; see bath_oracle.h).
 *
 * Easel pieces restated from the published Easel library (the reference needs
 * TravisWheelerLab/easel branch BATH, which is not vendored in /root/reference):
 *   esl_alphabet.c  amino "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~", DNA "ACGT-RYMKSWHBVDN*~"
 *   esl_gencode.c   NCBI translation tables, basic[16*n1+4*n2+n3] with ACGT order
 *   esl_sse.c       esl_sse_expf (Cephes single-precision exp, one lane)
 * Call sites pinning them: src/modelconfig.c:349,364-365; src/impl_sse/p7_fs_oprofile.c:252;
 * src/impl_sse/null2_fs.c:133.
 */
#include <string.h>
#include <math.h>
#include "bath_oracle.h"

static const char AA_SYMS[] = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~";
static const char NT_SYMS[] = "ACGT-RYMKSWHBVDN*~";

int bo_aa_digitize(char c)
{
  if (c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');
  if (c == '_' || c == '.') c = '-';
  const char *p = strchr(AA_SYMS, c);
  if (p == NULL || c == '\0') return -1;
  return (int)(p - AA_SYMS);
}

char bo_aa_symbol(int x) { return (x >= 0 && x < BO_KP) ? AA_SYMS[x] : '?'; }

int bo_nt_digitize(char c)
{
  if (c >= 'a' && c <= 'z') c = (char)(c - 'a' + 'A');
  if (c == 'U') c = 'T';
  if (c == 'X') c = 'N';
  if (c == 'I') c = 'A';   /* Easel maps inosine to A */
  if (c == '_' || c == '.') c = '-';
  const char *p = strchr(NT_SYMS, c);
  if (p == NULL || c == '\0') return -1;
  return (int)(p - NT_SYMS);
}

/* Easel esl_sq_ReverseComplement on digital DNA: complement table over
 * ACGT-RYMKSWHBVDN*~  ->  TGCA-YRKMSWDVBHN*~ */
void bo_dna_revcomp(uint8_t *dsq, int64_t L)
{
  static const uint8_t comp[18] = { 3, 2, 1, 0, 4, 6, 5, 8, 7, 9, 10, 14, 13, 12, 11, 15, 16, 17 };
  int64_t i, j;
  for (i = 1, j = L; i <= j; i++, j--) {
    uint8_t a = dsq[i], b = dsq[j];
    dsq[i] = (b < 18) ? comp[b] : b;
    dsq[j] = (a < 18) ? comp[a] : a;
  }
}

/* NCBI genetic codes.  Standard code in ACGT order (esl_gencode.c table 1): */
static const char STD_CODE[] = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF";

static int codon_index(const char *c3)
{
  int n1 = bo_nt_digitize(c3[0]), n2 = bo_nt_digitize(c3[1]), n3 = bo_nt_digitize(c3[2]);
  return 16 * n1 + 4 * n2 + n3;
}

const uint8_t *bo_gencode_basic(int ct)
{
  static uint8_t tbl[34][64];
  static int     built[34];
  char code[65];
  int  i;

  if (ct < 1 || ct > 33) return NULL;
  memcpy(code, STD_CODE, 65);
  switch (ct) {
  case 1: case 11: break;
  case 2:  code[codon_index("AGA")] = '*'; code[codon_index("AGG")] = '*';
           code[codon_index("ATA")] = 'M'; code[codon_index("TGA")] = 'W'; break;
  case 3:  code[codon_index("ATA")] = 'M'; code[codon_index("CTT")] = 'T'; code[codon_index("CTC")] = 'T';
           code[codon_index("CTA")] = 'T'; code[codon_index("CTG")] = 'T'; code[codon_index("TGA")] = 'W'; break;
  case 4:  code[codon_index("TGA")] = 'W'; break;
  case 5:  code[codon_index("AGA")] = 'S'; code[codon_index("AGG")] = 'S';
           code[codon_index("ATA")] = 'M'; code[codon_index("TGA")] = 'W'; break;
  case 6:  code[codon_index("TAA")] = 'Q'; code[codon_index("TAG")] = 'Q'; break;
  case 9:  code[codon_index("AAA")] = 'N'; code[codon_index("AGA")] = 'S';
           code[codon_index("AGG")] = 'S'; code[codon_index("TGA")] = 'W'; break;
  case 10: code[codon_index("TGA")] = 'C'; break;
  case 12: code[codon_index("CTG")] = 'S'; break;
  case 13: code[codon_index("AGA")] = 'G'; code[codon_index("AGG")] = 'G';
           code[codon_index("ATA")] = 'M'; code[codon_index("TGA")] = 'W'; break;
  case 14: code[codon_index("AAA")] = 'N'; code[codon_index("AGA")] = 'S'; code[codon_index("AGG")] = 'S';
           code[codon_index("TAA")] = 'Y'; code[codon_index("TGA")] = 'W'; break;
  case 16: code[codon_index("TAG")] = 'L'; break;
  case 21: code[codon_index("TGA")] = 'W'; code[codon_index("ATA")] = 'M'; code[codon_index("AGA")] = 'S';
           code[codon_index("AGG")] = 'S'; code[codon_index("AAA")] = 'N'; break;
  case 22: code[codon_index("TCA")] = '*'; code[codon_index("TAG")] = 'L'; break;
  case 23: code[codon_index("TTA")] = '*'; break;
  case 24: code[codon_index("AGA")] = 'S'; code[codon_index("AGG")] = 'K'; code[codon_index("TGA")] = 'W'; break;
  case 25: code[codon_index("TGA")] = 'G'; break;
  default: return NULL;
  }
  if (!built[ct]) {
    for (i = 0; i < 64; i++) tbl[ct][i] = (uint8_t) bo_aa_digitize(code[i]);
    built[ct] = 1;
  }
  return tbl[ct];
}

/* amino degeneracies (esl_alphabet.c set_amino): B={D,N} J={I,L} Z={E,Q} O={K} U={C} X=all */
static int aa_degen(int x, int y)
{
  switch (x) {
  case 21: return (y == 2  || y == 11);  /* B: D,N */
  case 22: return (y == 7  || y == 9);   /* J: I,L */
  case 23: return (y == 3  || y == 13);  /* Z: E,Q */
  case 24: return (y == 8);              /* O: K   */
  case 25: return (y == 1);              /* U: C   */
  case 26: return 1;                     /* X      */
  default: return 0;
  }
}

/* esl_abc_FExpectScVec(): expected score of degenerate residues, float accumulators */
void bo_abc_FExpectScVec(float *sc, const float *p)
{
  int x, i;
  for (x = BO_K + 1; x <= BO_KP - 3; x++) {
    float result = 0.0f, denom = 0.0f;
    for (i = 0; i < BO_K; i++)
      if (aa_degen(x, i)) { result += sc[i] * p[i]; denom += p[i]; }
    sc[x] = result / denom;
  }
}

/* esl_abc_FAvgScVec(): average score of degenerate residues */
void bo_abc_FAvgScVec(float *sc)
{
  int x, i;
  for (x = BO_K + 1; x <= BO_KP - 3; x++) {
    float result = 0.0f; int n = 0;
    for (i = 0; i < BO_K; i++)
      if (aa_degen(x, i)) { result += sc[i]; n++; }
    sc[x] = result / (float) n;
  }
}

/* esl_sse_expf(), one lane: Cephes expf with range reduction k = floor(x/ln2 + 0.5),
 * degree-5 polynomial on the remainder, 2^k built as an IEEE754 exponent. */
float bo_cephes_expf(float x)
{
  static const float cephes_p[6] = { 1.9875691500E-4f, 1.3981999507E-3f, 8.3334519073E-3f,
                                     4.1665795894E-2f, 1.6666665459E-1f, 5.0000001201E-1f };
  static const float cephes_c[2] = { 0.693359375f, -2.12194440e-4f };
  static const float maxlogf = 88.72283905206835f;
  static const float minlogf = -103.27892990343185f;
  float fx, tmp, z, y;
  int   k;
  union { int32_t i; float f; } u;

  if (x > maxlogf)  return INFINITY;
  if (x <= minlogf) return 0.0f;
  if (x != x)       return x;

  fx  = x * 1.44269504088896341f;   /* eslCONST_LOG2R */
  fx  = fx + 0.5f;
  k   = (int) fx;                   /* truncation */
  tmp = (float) k;
  if (tmp > fx) tmp -= 1.0f;        /* floor */
  fx  = tmp;
  k   = (int) fx;

  tmp = fx * cephes_c[0];
  z   = fx * cephes_c[1];
  x   = x - tmp;
  x   = x - z;
  z   = x * x;

  y = cephes_p[0];  y = y * x;
  y = y + cephes_p[1]; y = y * x;
  y = y + cephes_p[2]; y = y * x;
  y = y + cephes_p[3]; y = y * x;
  y = y + cephes_p[4]; y = y * x;
  y = y + cephes_p[5]; y = y * z;
  y = y + x;
  y = y + 1.0f;

  if (k + 127 <= 0) return 0.0f;    /* flush-to-zero regime (impl_Init sets FTZ/DAZ) */
  u.i = (k + 127) << 23;
  return y * u.f;
}
