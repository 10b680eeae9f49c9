/* bathhost.h -- C ABI of libbathhost.so: the HOST side of the translated-search path.
 *
 * In a GPU build of bathsearch these jobs stay in the reference's own C code (it has Easel);
 * this library restates them in C++ so that the path can be driven and measured end to end
 * without Easel: profile file reading, null model, frameshift profile construction and its
 * odds-ratio form, the length models, and (pipeline.cpp) the stage-batched translated pipeline.
 * Everything here runs once per query or per hit; the DP runs in libbathgpu.so (bathgpu.h).
 *
 *   bathhost_model_read       <- p7_hmmfile_Read (BATH3/f ASCII)              src/p7_hmmfile.c:1374-1690
 *                                + p7_bg_Create                                src/p7_bg.c:52-82
 *                                + p7_ProfileConfig_fs (3 and 5 codon lengths) src/modelconfig.c:220-698
 *                                + p7_fs_oprofile_Convert                      src/impl_sse/p7_fs_oprofile.c:222-296
 *                                as bathsearch sets a query up                 src/bathsearch.c:794-801
 *   bathhost_length_model     <- p7_fs_oprofile_ReconfigLength                 src/impl_sse/p7_fs_oprofile.c:636-651
 */
#ifndef BATHHOST_H
#define BATHHOST_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BATHHOST_OK       0
#define BATHHOST_EFAIL    1
#define BATHHOST_EOF      3
#define BATHHOST_EMEM     5
#define BATHHOST_EFORMAT  7
#define BATHHOST_EINVAL  11

typedef struct bathhost_model bathhost_model;

/* evparam order: MMU, MLAMBDA, VMU, VLAMBDA, FTAU, FLAMBDA, FTAUFS3, FTAUFS5 (src/hmmer.h:67) */
typedef struct {
  int32_t M;
  int32_t max_length;     /* MAXL, amino units; -1 if absent */
  int32_t codon_table;    /* CODON TABLE; -1 if absent */
  float   fsprob;         /* FRAMESHIFT PROB; -1 if absent */
  float   evparam[8];
  int32_t has_fs3_stats, has_fs5_stats;
  char    name[128];
  char    acc[64];
} bathhost_model_info;

/* Reads the index-th model of a .bhmm file and configures it the way bathsearch does for a
 * query (p7_LOCAL, dummy L=100).  ct <= 0: use the file's CODON TABLE (1 if absent). */
int  bathhost_model_read(const char *path, int index, int ct, bathhost_model **ret_model);
int  bathhost_model_count(const char *path);
void bathhost_model_destroy(bathhost_model *m);
int  bathhost_model_get_info(const bathhost_model *m, bathhost_model_info *info);

/* Un-striped odds-ratio tables in the layout bathgpu_load_fs_profile takes.
 * which = 3 | 5.  rfv: [nrows][M+1], tfv: [8][M+1] (BM,MM,IM,DM,MD,MI,II,DD; source-node indexed). */
int          bathhost_model_nrows(const bathhost_model *m, int which);
const float *bathhost_model_rfv(const bathhost_model *m, int which);
const float *bathhost_model_tfv(const bathhost_model *m, int which);
/* best amino acid / indel pattern per (node, codon row): P7_FS_PROFILE codons[][] / indel_pos[][]
 * (src/hmmer.h:372-411), [(M+1)][maxcodons] */
const uint8_t *bathhost_model_codons(const bathhost_model *m, int which);
const uint8_t *bathhost_model_indel_pos(const bathhost_model *m, int which);
/* core-model match emission probabilities [(M+1)][20] (row 0 unused) and consensus [M+2] */
const float *bathhost_model_mat(const bathhost_model *m);
const char  *bathhost_model_consensus(const bathhost_model *m);

/* N/C/J move and loop odds for a target of L_amino residues with nj expected J uses
 * (multihit local: nj = 1; unihit: nj = 0), in float as the reference computes them. */
void bathhost_length_model(int L_amino, float nj, float *pmove, float *ploop);

#ifdef __cplusplus
}
#endif
#endif
