/* bath_oracle.h -- CPU ORACLE for the BATH translated-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (TravisWheelerLab/BATH, src/ and src/impl_sse/), written in k-order
 * scalar loops.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may build, link or call it.  The product (bath_b200/, libbathgpu.so) never does.
 *
 * Parity status: the reference cannot be compiled here (its Easel dependency is
 * not vendored), so the oracle is pinned against the golden outputs the reference
 * ships: tutorial/AMP_N-fs.out|.tbl (see tests/test_oracle_golden.py).
 * Pieces of Easel used on the path (esl_sse_expf, esl_abc_FExpectScVec,
 * esl_gencode, esl_hmm Forward, gumbel/exponential tails) are restated from the
 * published Easel algorithms; each such function says so.
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef BATH_ORACLE_H
#define BATH_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Easel status codes (public Easel easel.h) */
#define BO_OK        0
#define BO_FAIL      1
#define BO_EOF       3
#define BO_EMEM      5
#define BO_EFORMAT   7
#define BO_EINVAL   11
#define BO_ERANGE   16
#define BO_ENORESULT 19

/* alphabets (Easel digital codes) */
#define BO_K    20
#define BO_KP   29
#define BO_AA_X     26   /* Kp-3 */
#define BO_AA_STOP  27   /* Kp-2 '*' */
#define BO_DSQ_SENTINEL 255
#define BO_MAXNUC 4

/* hmmer.h:125-131 */
enum { BO_H_MM = 0, BO_H_MI, BO_H_MD, BO_H_IM, BO_H_II, BO_H_DM, BO_H_DD };
/* hmmer.h:221-231 */
enum { BO_P_MM = 0, BO_P_IM, BO_P_DM, BO_P_BM, BO_P_MD, BO_P_DD, BO_P_MI, BO_P_II };
#define BO_P_NTRANS 8
/* hmmer.h:202-215 */
enum { BO_X_E = 0, BO_X_N, BO_X_J, BO_X_C };
enum { BO_X_LOOP = 0, BO_X_MOVE = 1 };
/* impl_sse.h:71-73 (optimized profile: MOVE=0, LOOP=1) */
enum { BO_O_MOVE = 0, BO_O_LOOP = 1 };
/* our unstriped transition block order (source-node indexed) */
enum { BO_T_BM = 0, BO_T_MM, BO_T_IM, BO_T_DM, BO_T_MD, BO_T_MI, BO_T_II, BO_T_DD };
/* hmmer.h:67 */
enum { BO_MMU = 0, BO_MLAMBDA, BO_VMU, BO_VLAMBDA, BO_FTAU, BO_FLAMBDA, BO_FTAUFS3, BO_FTAUFS5 };
/* hmmer.h:55-59 */
enum { BO_NO_MODE = 0, BO_LOCAL = 1, BO_GLOCAL = 2, BO_UNILOCAL = 3, BO_UNIGLOCAL = 4 };

/* hmmer.h:282-300 */
#define BO_MAXCODONS5 1367
#define BO_MAXCODONS3 338
#define BO_DEGEN5_C   1364
#define BO_DEGEN5_QC1 1365
#define BO_DEGEN5_QC2 1366
#define BO_DEGEN3_C   336
#define BO_DEGEN3_QC1 337

/* hmmer.h:306-314 (p7P_C2=1, C3=2, C4=3, C5=4) */
#define BO_CODON1_FS5(x)          ((x) * 341)
#define BO_CODON2_FS5(w,x)        ((x) * 341 + (w) * 85 + 1)
#define BO_CODON3_FS5(v,w,x)      ((x) * 341 + (w) * 85 + (v) * 21 + 2)
#define BO_CODON4_FS5(u,v,w,x)    ((x) * 341 + (w) * 85 + (v) * 21 + (u) * 5 + 3)
#define BO_CODON5_FS5(t,u,v,w,x)  ((x) * 341 + (w) * 85 + (v) * 21 + (u) * 5 + (t) + 4)
#define BO_CODON2_FS3(w,x)        ((x) * 84 + (w) * 21)
#define BO_CODON3_FS3(v,w,x)      ((x) * 84 + (w) * 21 + (v) * 5 + 1)
#define BO_CODON4_FS3(u,v,w,x)    ((x) * 84 + (w) * 21 + (v) * 5 + (u) + 2)
#define BO_MINIDX(a,b)            (((a) < (b)) ? (a) : (b))

/* hmmer.h:251-268 indel patterns */
enum { BO___X = 0, BO_X__, BO_XX_, BO_X_X, BO__XX, BO_XXX, BO_XXx, BO_XxX, BO_xXX, BO_xxx,
       BO_XXxX, BO_XxXX, BO_xXXX, BO_XXxxX, BO_XxxXX, BO_xxXXX };

/* impl_sse.h:324 X cells */
enum { BO_XC_E = 0, BO_XC_N, BO_XC_J, BO_XC_B, BO_XC_C, BO_XC_SCALE };
#define BO_NXCELLS 6
/* impl_sse.h:296-314: full FS matrix cells per (i,k): D, I, M_C0..M_C5 */
enum { BO_FS_D = 0, BO_FS_I = 1, BO_FS_M = 2 };
#define BO_NSCELLS_FS 8
/* bck / OA matrices: M, D, I */
enum { BO_S_M = 0, BO_S_D = 1, BO_S_I = 2 };
#define BO_NSCELLS 3

/* trace states (hmmer.h p7t_statetype_e) */
enum { BO_T_BOGUS = 0, BO_ST_M = 1, BO_ST_D = 2, BO_ST_I = 3, BO_ST_S = 4, BO_ST_N = 5,
       BO_ST_B = 6, BO_ST_E = 7, BO_ST_C = 8, BO_ST_T = 9, BO_ST_J = 10, BO_ST_X = 11 };

/* ---- core HMM (p7_hmmfile.c:1374-1690) ---- */
typedef struct {
  int    M;
  int    max_length;
  char   name[128];
  char   acc[64];
  float  evparam[8];
  int    has_stats_fs3, has_stats_fs5;
  float  fsprob;
  int    ct;
  float  compo[BO_K];
  int    has_compo;
  float *t;     /* [(M+1)][7]  */
  float *mat;   /* [(M+1)][20] */
  float *ins;   /* [(M+1)][20] */
  char  *consensus; /* [M+2], 1..M */
} BO_HMM;

/* ---- null model (p7_bg.c) ---- */
typedef struct {
  float f[BO_K];
  float p1;
  float omega;
  /* 2-state bias filter HMM (p7_bg.c:449-471; Easel esl_hmm) */
  float fh_t[2][3];     /* t[k][0..1] transitions, t[k][2] = end */
  float fh_e[2][BO_K];  /* emission probabilities */
  float fh_eo[2][BO_KP];/* emission odds ratios (esl_hmm_Configure) */
  float fh_pi[3];
} BO_BG;

/* ---- generic protein profile, log-odds (modelconfig.c:48-196) ---- */
typedef struct {
  int    M, L, mode, max_length;
  float  nj;
  float *tsc;   /* [M][8]   nodes 0..M-1 */
  float *rsc;   /* [Kp][(M+1)*2] */
  float  xsc[4][2];
  float  evparam[8];
  float  compo[BO_K];
} BO_PROFILE;

/* ---- generic frameshift profile, log-odds (modelconfig.c:220-698) ---- */
typedef struct {
  int      M, L, mode, max_length, codon_lengths, maxcodons;
  float    nj;
  float    fsprob;
  float   *tsc;        /* [M][8] */
  float   *rsc;        /* [(maxcodons+Kp)][M+1] */
  float    xsc[4][2];
  uint8_t *codons;     /* [(M+1)][maxcodons] best amino acid  */
  uint8_t *indel_pos;  /* [(M+1)][maxcodons] indel pattern    */
  float    evparam[8];
} BO_FS_PROFILE;

/* ---- "optimized" frameshift profile in odds-ratio space, UN-striped
 *      (p7_fs_oprofile.c:222-296 without the SSE striping) ---- */
typedef struct {
  int    M, L, mode, codon_lengths, maxcodons, nrows;
  float  nj;
  float *rfv;     /* [nrows][M+1] odds; column 0 is 0.0 */
  float *tfv;     /* [8][M+1]  tfv[BO_T_x][k] = exp(tsc[k][x]) for k=0..M-1, 0.0 at k=M */
  float  xf[4][2];/* [E,N,J,C][MOVE,LOOP] */
  float  evparam[8];
} BO_FS_OPROFILE;

/* ---- "optimized" protein profile, UN-striped (impl_sse.h:75-142): byte costs for MSV/SSV, word scores
 *      for the Viterbi filter, float odds ratios for Forward/Backward ---- */
typedef struct {
  int      M, L, mode, max_length;
  float    nj;
  /* MSV/SSV: uint8 costs */
  uint8_t *rbv;        /* [Kp][M+1] biased match costs, column 0 = 255 */
  uint8_t  tbm_b, tec_b, tjb_b, base_b, bias_b;
  float    scale_b;
  /* Viterbi filter: int16 scores */
  int16_t *rwv;        /* [Kp][M+1] */
  int16_t *twv;        /* [8][M+1] BM,MM,IM,DM,MD,MI,II,DD, source-node indexed */
  int16_t  xw[4][2];   /* [E,N,J,C][MOVE,LOOP] */
  int16_t  base_w, ddbound_w;
  float    scale_w;
  /* Forward/Backward: float odds ratios */
  float   *rfv;        /* [Kp][M+1] */
  float   *tfv;        /* [8][M+1] */
  float    xf[4][2];
  float    evparam[8];
  float    compo[BO_K];
} BO_OPROFILE;

/* P7_HMM_WINDOW / P7_HMM_WINDOWLIST (src/hmmer.h, p7_hmmwindow.c:83) */
typedef struct { int n, k, length, target_len, id; float score; } BO_WINDOW;
typedef struct { BO_WINDOW *w; int count, nalloc; } BO_WINDOWLIST;

/* ---- DP matrices ---- */
typedef struct {
  int    M, L;
  int    allocL;
  int    nscells;     /* 0 (parser: no MDI kept), 3 or 8 */
  float *dp;          /* [(L+1)][(M+1)][nscells] if nscells>0 */
  float *xmx;         /* [(L+1)][6] */
  float  totscale;
  int    has_own_scales;
} BO_MX;

/* frameshift trace (p7_trace.c fs variants) */
typedef struct {
  int    N, nalloc;
  int    M, L;
  char  *st;
  int   *k;
  int   *i;
  int   *c;
  float *pp;
} BO_TRACE;

/* ===== alphabet.c ===== */
int   bo_aa_digitize(char c);
int   bo_nt_digitize(char c);
char  bo_aa_symbol(int x);
void  bo_dna_revcomp(uint8_t *dsq, int64_t L);         /* dsq 1..L in place */
const uint8_t *bo_gencode_basic(int ct);               /* [64] amino codes, stop=27; NULL if unsupported */
void  bo_abc_FExpectScVec(float *sc, const float *p);  /* Easel esl_abc_FExpectScVec, amino */
void  bo_abc_FAvgScVec(float *sc);                     /* Easel esl_abc_FAvgScVec, amino */
float bo_cephes_expf(float x);                         /* Easel esl_sse_expf, one lane */

/* ===== logsum.c ===== */
void  bo_FLogsumInit(void);
float bo_FLogsum(float a, float b);

/* ===== hmmfile.c ===== */
int   bo_hmmfile_read(const char *path, int index, BO_HMM **ret_hmm);  /* index-th model in file */
int   bo_hmmfile_count(const char *path);
void  bo_hmm_destroy(BO_HMM *hmm);

/* ===== profile.c ===== */
BO_BG *bo_bg_create(void);
void   bo_bg_destroy(BO_BG *bg);
void   bo_bg_SetLength(BO_BG *bg, int L);
float  bo_bg_NullOne(const BO_BG *bg, int L);
float  bo_bg_fs_NullOne(const BO_BG *bg, int aminoL);
void   bo_hmm_CalculateOccupancy(const BO_HMM *hmm, float *mocc);
BO_PROFILE *bo_profile_config(const BO_HMM *hmm, const BO_BG *bg, int L, int mode);
void   bo_profile_destroy(BO_PROFILE *gm);
void   bo_profile_ReconfigLength(BO_PROFILE *gm, int L);
BO_FS_PROFILE *bo_fs_profile_config(const BO_HMM *hmm, const BO_BG *bg, int ct, int codon_lengths, int L_amino, int mode);
void   bo_fs_profile_destroy(BO_FS_PROFILE *gm);
void   bo_fs_ReconfigLength(BO_FS_PROFILE *gm, int L_amino);
void   bo_fs_ReconfigUnihit(BO_FS_PROFILE *gm, int L_amino);
void   bo_fs_ReconfigMultihit(BO_FS_PROFILE *gm, int L_amino);
BO_FS_OPROFILE *bo_fs_oprofile_convert(const BO_FS_PROFILE *gm);
void   bo_fs_oprofile_destroy(BO_FS_OPROFILE *om);
void   bo_fs_oprofile_ReconfigLength(BO_FS_OPROFILE *om, int L);
void   bo_fs_oprofile_ReconfigUnihit(BO_FS_OPROFILE *om, int L);
void   bo_fs_oprofile_ReconfigMultihit(BO_FS_OPROFILE *om, int L);

/* ===== mx.c ===== */
BO_MX *bo_mx_create(int M, int L, int nscells);
void   bo_mx_destroy(BO_MX *mx);
BO_TRACE *bo_trace_create(void);
void   bo_trace_reuse(BO_TRACE *tr);
void   bo_trace_destroy(BO_TRACE *tr);
int    bo_trace_append(BO_TRACE *tr, char st, int k, int i, int c, float pp);
void   bo_trace_reverse(BO_TRACE *tr);

/* ===== fs_fwdback.c  (impl_sse/fwdback_fs.c) ===== */
int bo_ForwardParser_Frameshift_3Codons (const uint8_t *dsq, int L, const BO_FS_OPROFILE *om, BO_MX *ox, float *opt_sc);
int bo_BackwardParser_Frameshift_3Codons(const uint8_t *dsq, int L, const BO_FS_OPROFILE *om, const BO_MX *fwd, BO_MX *bck, float *opt_sc);
int bo_Forward_Frameshift (const uint8_t *dsq, int L, const BO_FS_OPROFILE *om, BO_MX *fwd, float *opt_sc);
int bo_Backward_Frameshift(const uint8_t *dsq, int L, const BO_FS_OPROFILE *om, const BO_MX *fwd, BO_MX *bck, float *opt_sc);

/* ===== fs_decoding.c (impl_sse/decoding_fs.c) ===== */
int bo_Decoding_Frameshift(const BO_FS_OPROFILE *om, BO_MX *fwd, const BO_MX *bck);
int bo_DomainDecoding_Frameshift(const float xf_loop_NJC[3], const BO_MX *oxf, const BO_MX *oxb,
                                 float *btot, float *etot, float *mocc);

/* ===== fs_optacc.c (impl_sse/optacc_fs.c) ===== */
int bo_OptimalAccuracy_Frameshift(const BO_FS_OPROFILE *om, const BO_MX *pp, BO_MX *ox, float *ret_e);
int bo_OATrace_Frameshift(const BO_FS_OPROFILE *om, const BO_MX *pp, const BO_MX *ox, BO_TRACE *tr);

/* ===== fs_stotrace.c (impl_sse/stotrace_fs.c, p7_spensemble.c, p7_domaindef.c:892-954) ===== */
typedef struct { uint32_t seed, x; } BO_RNG;                         /* Easel's fast generator (esl_randomness_CreateFast) */
typedef struct { int idx, i, j, k, m; float prob; } BO_SEGMENT;      /* struct p7_spcoord_s */
void   bo_rng_init(BO_RNG *r, uint32_t seed);
double bo_random(BO_RNG *r);
int bo_StochasticTrace_Frameshift(BO_RNG *rng, int L, const BO_FS_OPROFILE *om, const BO_MX *ox, BO_TRACE *tr);
int bo_trace_fs_Index(const BO_TRACE *tr, BO_SEGMENT *seg, int max_seg);
int bo_spensemble_fs_Cluster(const BO_SEGMENT *sp, int n, int nsamples, float min_overlap, int of_smaller, int max_diagdiff,
                             float min_posterior, float min_endpointp, BO_SEGMENT *out, int max_out);
int bo_region_trace_ensemble_frameshift(const BO_FS_OPROFILE *om, const BO_MX *fwd, int ireg, int jreg, uint32_t seed, int nsamples,
                                        BO_SEGMENT *samples, int max_samples, int *ret_nsamples, BO_SEGMENT *out, int max_out);
/* ===== stotrace.c: the standard-translation flavour (src/impl_sse/stotrace.c, null2.c:131-219, p7_domaindef.c:766-860) ===== */
int bo_spensemble_Cluster(const BO_SEGMENT *sp, int n, int nsamples, float min_overlap, int of_smaller, int max_diagdiff,
                          float min_posterior, float min_endpointp, BO_SEGMENT *out, int max_out);
int bo_StochasticTrace(BO_RNG *rng, int L, const BO_OPROFILE *om, const BO_MX *ox, BO_TRACE *tr);
int bo_region_trace_ensemble(const BO_OPROFILE *om, const uint8_t *dsq, const BO_MX *fwd, int ireg, int jreg, uint32_t seed, int nsamples,
                             BO_SEGMENT *samples, int max_samples, int *ret_nsamples, BO_SEGMENT *out, int max_out, float *n2sc);

/* ===== calibrate.c (src/evalues.c: p7_Calibrate, p7_Lambda, p7_MSVMu, p7_ViterbiMu, p7_Tau, p7_fs_Tau_3codons/_5codons) ===== */
double bo_Lambda(const BO_HMM *hmm, const BO_BG *bg);
double bo_gumbel_FitCompleteLoc(const double *x, int n, double lambda);
void   bo_gumbel_FitComplete(const double *x, int n, double *ret_mu, double *ret_lambda);

/* ===== fs_null2.c (impl_sse/null2_fs.c) ===== */
int bo_Null2_fs_ByExpectation(const BO_FS_OPROFILE *om, BO_MX *pp, float *null2 /* [Kp] */);

/* ===== filters.c (impl_sse/p7_oprofile.c, msvfilter.c, ssvfilter.c, vitfilter.c) ===== */
BO_OPROFILE *bo_oprofile_convert(const BO_PROFILE *gm);
void   bo_oprofile_destroy(BO_OPROFILE *om);
void   bo_oprofile_ReconfigLength(BO_OPROFILE *om, int L);
void   bo_oprofile_ssv_scores(const BO_OPROFILE *om, uint8_t *arr /* [(M+1)*Kp] */);
int    bo_SSVFilter(const uint8_t *dsq, int L, const BO_OPROFILE *om, float *ret_sc);
int    bo_MSVFilter(const uint8_t *dsq, int L, const BO_OPROFILE *om, float *ret_sc);
int    bo_MSVFilter_opt(const uint8_t *dsq, int L, const BO_OPROFILE *om, int use_ssv, float *ret_sc);
int    bo_SSVFilter_BATH(const uint8_t *dsq, int L, BO_OPROFILE *om, const uint8_t *ssv_scores, float nullsc, double P,
                         int lanes, BO_WINDOWLIST *wl);
int    bo_ViterbiFilter(const uint8_t *dsq, int L, const BO_OPROFILE *om, float *ret_sc);
int    bo_ViterbiFilter_BATH(const uint8_t *dsq, int L, const BO_OPROFILE *om, const uint8_t *ssv_scores, float filtersc, double P,
                             int lanes, BO_WINDOWLIST *wl, float *ret_sc);
int    bo_SSVFilter_BATH_thresh(const uint8_t *dsq, int L, const BO_OPROFILE *om, const uint8_t *ssv_scores, uint8_t sc_thresh,
                                int lanes, BO_WINDOWLIST *wl);
int    bo_ViterbiFilter_BATH_thresh(const uint8_t *dsq, int L, const BO_OPROFILE *om, const uint8_t *ssv_scores, int16_t sc_thresh,
                                    int sc_ext_thresh, int lanes, BO_WINDOWLIST *wl, float *ret_sc);
void   bo_windowlist_reset(BO_WINDOWLIST *wl);
void   bo_windowlist_free(BO_WINDOWLIST *wl);
double bo_gumbel_invsurv(double p, double mu, double lambda);
double bo_gumbel_surv(double x, double mu, double lambda);
double bo_exp_surv(double x, double mu, double lambda);
double bo_exp_logsurv(double x, double mu, double lambda);

/* ===== orf_fwd.c (impl_sse/fwdback.c) ===== */
int    bo_ForwardParser(const uint8_t *dsq, int L, const BO_OPROFILE *om, float *opt_sc);

/* ===== orfs.c: six-frame translation (Easel esl_gencode semantics) ===== */
typedef struct { int start, end, frame, n; int64_t offset; } BO_ORF;    /* nucleotide coordinates on the oriented strand, start < end */
int    bo_find_orfs(const uint8_t *dsq, int n, const uint8_t *gcode, int min_len, BO_ORF **ret_orfs, int *ret_n, uint8_t **ret_res, int64_t *ret_nres);

/* ===== orf_domain.c: the standard-translation branch's DP over an ORF (src/impl_sse/fwdback.c, decoding.c, optacc.c, null2.c) ===== */
void   bo_oprofile_ReconfigMultihit(BO_OPROFILE *om, int L);
void   bo_oprofile_ReconfigUnihit(BO_OPROFILE *om, int L);
int    bo_Forward (const uint8_t *dsq, int L, const BO_OPROFILE *om, BO_MX *ox, float *opt_sc);     /* ox->nscells 3: full; 0: parser */
int    bo_Backward(const uint8_t *dsq, int L, const BO_OPROFILE *om, const BO_MX *fwd, BO_MX *bck, float *opt_sc);
int    bo_Decoding(const BO_OPROFILE *om, const BO_MX *oxf, BO_MX *oxb, BO_MX *pp);
int    bo_DomainDecoding(const float xf_loop_NJC[3], const BO_MX *oxf, const BO_MX *oxb, int own_scales, float *btot, float *etot, float *mocc);
int    bo_OptimalAccuracy(const BO_OPROFILE *om, const BO_MX *pp, BO_MX *ox, float *ret_e);
int    bo_OATrace(const BO_OPROFILE *om, const BO_MX *pp, const BO_MX *ox, int lanes, BO_TRACE *tr);
int    bo_Null2_ByExpectation(const BO_OPROFILE *om, const BO_MX *pp, float *null2);

/* ===== batch.c (worker-thread pool over windows; src/bathsearch.c:814-844,1224) ===== */
int bo_batch_ForwardParser_3Codons(const uint8_t *dsq, const int64_t *start, const int32_t *L, int n,
                                   const BO_FS_OPROFILE *om, int nthreads, float *sc, int32_t *status);
/* ===== fwd3_avx2.c: the same batch on an AVX2 + FMA build of the parser (the CPU arm of bench.py; checked against the scalar one) ===== */
int bo_fwd3_simd_supported(void);
/* msv_avx2.c: AVX2 build of p7_MSVFilter + SSV shortcut for the CPU arm of bench.py's search metric (bit-identical to bo_MSVFilter) */
typedef struct bo_msv_simd_s bo_msv_simd;
int          bo_msv_simd_supported(void);
bo_msv_simd *bo_msv_simd_create(const BO_OPROFILE *om);
void         bo_msv_simd_destroy(bo_msv_simd *im);
int          bo_MSVFilter_simd(const bo_msv_simd *im, const uint8_t *dsq, int L, const BO_OPROFILE *om, float *ret_sc);
int bo_batch_ForwardParser_3Codons_simd(const uint8_t *dsq, const int64_t *start, const int32_t *L, int n,
                                        const BO_FS_OPROFILE *om, int nthreads, float *sc, int32_t *status);

#ifdef __cplusplus
}
#endif
int bo_Calibrate(const BO_HMM *hmm, BO_BG *bg, BO_OPROFILE *om, BO_FS_OPROFILE *om_fs3, BO_FS_OPROFILE *om_fs5, int ct,
                 uint32_t seed, double lambda, int which_mask, int convert_flow, uint32_t *rng_x, double out[8]);

#endif
