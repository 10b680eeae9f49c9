// host_internal.h -- types shared by the host-side sources of libbathhost.so.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace bathhost {

constexpr int kK  = 20;   // canonical amino acids
constexpr int kKp = 29;   // all amino codes: 20 + gap + BJZOUX + '*' + '~'

// core-model transition order (src/hmmer.h:125-131) and profile transition order (:221-231)
enum { HT_MM = 0, HT_MI, HT_MD, HT_IM, HT_II, HT_DM, HT_DD };
enum { PT_MM = 0, PT_IM, PT_DM, PT_BM, PT_MD, PT_DD, PT_MI, PT_II };
enum { PX_E = 0, PX_N, PX_J, PX_C };
enum { PX_LOOP = 0, PX_MOVE = 1 };
enum { EV_MMU = 0, EV_MLAMBDA, EV_VMU, EV_VLAMBDA, EV_FTAU, EV_FLAMBDA, EV_FTAUFS3, EV_FTAUFS5 };

struct CoreModel {                 // P7_HMM as read from a BATH3/f file (probabilities)
  int   M = 0, max_length = -1, ct = -1;
  float fsprob = -1.0f;
  float evparam[8] = { -99999.f, -99999.f, -99999.f, -99999.f, -99999.f, -99999.f, -99999.f, -99999.f };
  bool  has_fs3 = false, has_fs5 = false, has_compo = false;
  float compo[kK] = { 0 };
  std::string name, acc, desc;
  std::string rf, cs;              // [M+2] reference / consensus-structure annotation, 1..M; empty if the file has none
  std::vector<float> t, mat, ins;  // [(M+1)][7], [(M+1)][20], [(M+1)][20]
  std::vector<char>  consensus;    // [M+2], 1..M
};

struct NullModel {                 // P7_BG
  float f[kK];
  float p1, omega;
  NullModel();
};

struct FsProfile {                 // P7_FS_PROFILE (log-odds)
  int   M = 0, codon_lengths = 0, maxcodons = 0;
  float nj = 1.0f;
  float xsc[4][2] = { { 0 } };
  std::vector<float>   rsc;        // [(maxcodons+Kp)][M+1]
  std::vector<float>   tsc;        // [M][8]
  std::vector<uint8_t> codons, indel_pos;   // [(M+1)][maxcodons]
  void configure(const CoreModel &h, const NullModel &bg, const uint8_t gcode[64], int codon_lengths);
};

struct FsOddsProfile {             // P7_FS_OPROFILE without the SIMD striping
  int   M = 0, codon_lengths = 0, nrows = 0;
  float xfE_move = 0.5f, xfE_loop = 0.5f;
  std::vector<float> rfv;          // [nrows][M+1]
  std::vector<float> tfv;          // [8][M+1]
  void convert(const FsProfile &gm);
};

// P7_OPROFILE's integer and float parts for the ORF stage, un-striped (src/impl_sse/p7_oprofile.c:667-985),
// plus the P7_SCOREDATA pieces the pipeline reads (src/p7_scoredata.c:60-70, :296-375)
struct ProteinProfile {
  int   M = 0, max_length = -1;
  float nj = 1.0f;
  // generic profile (src/modelconfig.c:48-196): match log-odds and the specials that never change
  std::vector<float> msc;          // [Kp][M+1]
  float xsc_E_move = 0, xsc_E_loop = 0;
  // bytes
  float scale_b = 0;
  int   base_b = 190, bias_b = 0, tbm_b = 0, tec_b = 0;
  std::vector<uint8_t> rbv;        // [Kp][M+1]
  // words
  float scale_w = 0;
  int   base_w = 12000, ddbound_w = -32768, xw_E_move = 0, xw_E_loop = 0;
  std::vector<int16_t> rwv;        // [Kp][M+1]
  std::vector<int16_t> twv;        // [8][M+1]
  // score data
  std::vector<float> prefix_lengths, suffix_lengths;   // [M+1]
  void configure(const CoreModel &h, const NullModel &bg, const FsProfile &gm_fs, const FsOddsProfile &om_fs);
  uint8_t unbiased_byteify(float sc) const;
  uint8_t biased_byteify(float sc) const;
  int16_t wordify(float sc) const;
  uint8_t tjb_for_length(int L) const;      // p7_oprofile_ReconfigMSVLength
  int16_t xw_move_for_length(int L) const;  // p7_oprofile_ReconfigRestLength
};

// ---- stotrace.cpp: the multi-domain branch of frameshift domain definition
struct Segment { int idx, i, j, k, m; float prob; };             // struct p7_spcoord_s (src/p7_spensemble.c)
struct ForwardMatrix {                                            // one region of bathgpu_fs_forward_matrices' output
  const float *mx;                                                // [(L+1)][(M+1)][8] {D, I, M_C0..M_C5}
  const float *xr;                                                // [(L+1)][6]        {E, N, J, B, C, SCALE}
  int M, L;
};
struct SpecialOdds { float move, loop, e_move, e_loop; };         // N/J/C -> MOVE and LOOP, E -> MOVE and LOOP of the multihit profile
bool sample_region_segments(const ForwardMatrix &F, const float *tfv, const SpecialOdds &X, uint32_t seed, int nsamples, int ireg,
                            std::vector<Segment> &out);
std::vector<Segment> cluster_region_segments(const std::vector<Segment> &sp, int nsamples, bool protein = false);
// the standard-translation flavour (region_trace_ensemble, src/p7_domaindef.c:766-860): mx = [(L+1)][(M+1)][4] {M, D, I, 0} of
// bathgpu_orf_forward_matrices, res[1..L] = the region's residues, rf = amino-acid emission odds [Kp][M+1]; besides the sampled
// segments (ORF coordinates: ireg = first residue of the region) n2sc[1..L] receives log(mean over the samples of the per-residue
// null2 odds), the position-specific null2 score rescore_isolated_domain_bath sums when null2_is_done
bool sample_region_segments_protein(const ForwardMatrix &F, const float *tfv, const float *rf, const SpecialOdds &X, uint32_t seed, int nsamples,
                                    int ireg, const uint8_t *res, std::vector<Segment> &out, std::vector<float> &n2sc);

int   builder_max_length(const CoreModel &h, double emit_thresh);
int   amino_code(char c);
int   dna_code(char c);
bool  genetic_code(int ct, uint8_t out[64]);
float simd_expf(float x);

}  // namespace bathhost
