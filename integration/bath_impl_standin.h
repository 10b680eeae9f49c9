/* bath_impl_standin.h -- a minimal STAND-IN for the declarations of BATH's impl layer that impl_cuda_shim.c touches.
 *
 * BATH cannot be built in this repository (Easel is not vendored, /root/reference/INSTALL:6-8), so the shim is compiled here against
 * this header instead of src/impl_sse/impl_sse.h + easel.h.  Every type mirrors the reference's field names, order and meaning for
 * the fields the shim reads or writes (reference file:line beside each); members the shim never touches are kept so that the struct
 * is recognisably the reference's, but their exact types do not matter here.  In a BATH tree the shim includes "hmmer.h" and
 * "impl_sse.h" instead (define BATH_SHIM_USE_REFERENCE_HEADERS) and this file is not used.
 */
#ifndef BATH_IMPL_STANDIN_H
#define BATH_IMPL_STANDIN_H
#include <stdint.h>
#include <stdio.h>
#include <sys/types.h>
#include <xmmintrin.h>
#include <emmintrin.h>

typedef uint8_t ESL_DSQ;                                   /* easel.h */
#define eslOK       0
#define eslEMEM     5
#define eslEINVAL  11
#define eslERANGE  16
#define eslINFINITY INFINITY
#define ESL_MAX(a, b) (((a) > (b)) ? (a) : (b))
typedef struct { int type, K, Kp; } ESL_ALPHABET;          /* esl_alphabet.h: only K / Kp are read */

#define p7_NOFFSETS  3                                     /* src/hmmer.h */
#define p7_NEVPARAM  8
#define p7_NCUTOFFS  6
#define p7_MAXABET   20
#define p7P_MAXCODONS5 1367                                /* src/hmmer.h:282 */
#define p7P_MAXCODONS3 338                                 /* src/hmmer.h:283 */

#define p7O_NQB(M)   (ESL_MAX(2, ((((M) - 1) / 16) + 1)))  /* src/impl_sse/impl_sse.h:24 */
#define p7O_NQW(M)   (ESL_MAX(2, ((((M) - 1) / 8) + 1)))   /* :25 */
#define p7O_NQF(M)   (ESL_MAX(2, ((((M) - 1) / 4) + 1)))   /* src/impl_sse/impl_sse.h:26 */
#define p7O_NXSTATES 4                                     /* :68 */
#define p7O_NXTRANS  2                                     /* :69 */
#define p7O_NTRANS   8                                     /* :70 */
enum p7o_xstates_e      { p7O_E = 0, p7O_N = 1, p7O_J = 2, p7O_C = 3 };                                             /* :71 */
enum p7o_xtransitions_e { p7O_MOVE = 0, p7O_LOOP = 1 };                                                             /* :72 */
enum p7o_tsc_e          { p7O_BM = 0, p7O_MM = 1, p7O_IM = 2, p7O_DM = 3, p7O_MD = 4, p7O_MI = 5, p7O_II = 6, p7O_DD = 7 };  /* :73 */

typedef struct p7_oprofile_s {                             /* src/impl_sse/impl_sse.h:74-139 */
  __m128i **rbv;                                           /* MSV match costs [x][q], 16 uchars per vector */
  __m128i **sbv;
  uint8_t   tbm_b, tec_b, tjb_b;
  float     scale_b;
  uint8_t   base_b, bias_b;
  __m128i **rwv;                                           /* Viterbi match scores [x][q], 8 words per vector */
  __m128i  *twv;                                           /* [8 * Q8] */
  int16_t   xw[p7O_NXSTATES][p7O_NXTRANS];
  float     scale_w;
  int16_t   base_w, ddbound_w;
  float     ncj_roundoff;
  __m128  **rfv;                                           /* Forward / Backward odds [x][q], 4 floats per vector */
  __m128   *tfv;                                           /* [8 * Q4] */
  float     xf[p7O_NXSTATES][p7O_NXTRANS];
  __m128i  *rbv_mem, *sbv_mem, *rwv_mem, *twv_mem;
  __m128   *tfv_mem, *rfv_mem;
  off_t     offs[p7_NOFFSETS];
  off_t     roff, eoff;
  char     *name, *acc, *desc, *rf, *mm, *cs, *consensus;
  float     evparam[p7_NEVPARAM];
  float     cutoff[p7_NCUTOFFS];
  float     compo[p7_MAXABET];
  const ESL_ALPHABET *abc;
  int       L, M, max_length, allocM, allocQ4, allocQ8, allocQ16, mode;
  float     nj;
  int       clone;
} P7_OPROFILE;

#define eslENORESULT 19                                    /* easel.h */

typedef struct p7_fs_oprofile_s {                          /* src/impl_sse/impl_sse.h:200-244 */
  __m128  **rfv;                                           /* [c][q], c = 0..p7P_MAXCODONS#+Kp-1, q = 0..allocQ4-1 */
  __m128   *tfv;                                           /* [p7O_NTRANS * allocQ4] */
  float     xf[p7O_NXSTATES][p7O_NXTRANS];
  int       codon_lengths;
  float     fsprob;
  __m128   *rfv_mem;
  __m128   *tfv_mem;
  off_t     offs[p7_NOFFSETS];
  off_t     roff;
  off_t     eoff;
  char     *name, *acc, *desc, *rf, *mm, *cs, *consensus;
  float     evparam[p7_NEVPARAM];
  float     cutoff[p7_NCUTOFFS];
  float     compo[p7_MAXABET];
  const ESL_ALPHABET *abc;
  int       L;
  int       M;
  int       max_length;
  int       allocM;
  int       allocQ4;
  int       mode;
  float     nj;
  int       clone;
} P7_FS_OPROFILE;

enum p7x_xcells_e { p7X_E = 0, p7X_N = 1, p7X_J = 2, p7X_B = 3, p7X_C = 4, p7X_SCALE = 5 };   /* :317 */
#define p7X_NXCELLS 6                                                                          /* :318 */

typedef struct p7_omx_s {                                  /* src/impl_sse/impl_sse.h:329-358 */
  int       M;
  int       L;
  int       nscells;
  __m128  **dpf;
  __m128i **dpw;
  __m128i **dpb;
  void     *dp_mem;
  int       allocR;
  int       validR;
  int       allocQ4;
  int       allocQ8;
  int       allocQ16;
  size_t    ncells;
  float    *xmx;                                           /* [i*p7X_NXCELLS + s], i = 0..L */
  void     *x_mem;
  int       allocXR;
  float     totscale;
  int       has_own_scales;
  int       debugging;
  FILE     *dfp;
} P7_OMX;

typedef struct p7_oivx_s P7_OIVX;                          /* :281 -- the CPU kernels' scratch rows; the shim ignores it */

#endif
