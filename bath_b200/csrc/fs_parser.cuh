// fs_parser.cuh -- frameshift Forward / Backward PARSER kernels (3 codon lengths) for sm_100a.
//
// What they compute is p7_ForwardParser_Frameshift_3Codons / p7_BackwardParser_Frameshift_3Codons
// (reference: src/impl_sse/fwdback_fs.c:97-533, :565-1013).  How they compute it is B200-first:
//
//  * one warp per DNA window; lane l owns J CONTIGUOUS model nodes k = l*J+1 .. l*J+J
//    (no striping -- striping exists to feed CPU SIMD lanes); all per-node state lives in registers;
//  * per-node state is the minimum the recurrence needs: the pre-emission entry values V(r)[k]
//    (the reference's IVX) for 4 rows and the insert values I(r)[k] for 3 rows -- 7 floats per node,
//    instead of the reference's 4x3 MDI ring + 3 IVX rows = 15;
//  * the D->D chain D(k+1) = M(k) tMD(k) + D(k) tDD(k) is a first-order linear recurrence: a serial
//    pass inside the lane plus a 5-step warp-shuffle scan whose multipliers are profile constants;
//  * E(i) = sum_k M(i,k)+D(i,k) is taken as a dot product sum_k M(i,k) Z(k) with the profile constant
//    Z(k) = 1 + tMD(k) (1 + tDD(k+1) + tDD(k+1) tDD(k+2) + ...), so E (hence B(i), hence row i+2)
//    does not wait for the D scan;
//  * the three emission rows of a DP row are read with coalesced 128/64/32-bit loads from a table laid
//    out [codon][J/VEC][lane][VEC]; the table is L1/L2 resident (0.3 MB at M=200);
//  * target nucleotides are 4-bit packed, fetched one 16-bit quad (4 rows) ahead;
//  * rows are unrolled by 4 so every ring slot is a compile-time register name; the window is entered
//    at row 0 and the loop may run up to 3 rows past L (their results are discarded).
//
// Rescaling follows the reference: when E(i) > 1e4 everything live is divided by E(i) and log E(i)
// is accumulated (fwdback_fs.c:472-496).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bathgpu {

constexpr int kWarp = 32;

// lane-constant image: [NC][J][32] floats then [NL][32] floats
enum FwdCellConst { FC_BM = 0, FC_MM, FC_IM, FC_DM, FC_MD, FC_DD, FC_PP, FC_MI, FC_II, FC_Z, FC_COUNT };
enum FwdLaneConst { FL_B0 = 0, FL_B1, FL_B2, FL_B3, FL_B4, FL_COUNT };

struct WindowDesc {      // device copy of bathgpu_window
  long long start;       // 1-based block coordinate of window position 1
  int       L;
  float     pmove;
  float     ploop;
};

struct FsParserArgs {
  const float    *emis;        // [nrows][mpad] permuted odds table (3-codon profile)
  const float    *cellc;       // lane-constant image
  const uint32_t *dna4;        // 4-bit packed block, nt p (0-based) at word (p+8)>>3 (one guard word in front)
  const WindowDesc *wins;
  int             nwin;
  int             mpad;        // 32*J
  float           tEM, tEL;    // E->MOVE, E->LOOP odds
  float          *fwdsc;       // [nwin]
  int            *status;      // [nwin]
  float          *xmx;         // optional: X rows, window w at xmx + xoff[w]*6
  const long long *xoff;
  int            *counter;     // work-queue counter
};

template <int J> struct VecOf { static constexpr int V = (J % 4 == 0) ? 4 : ((J % 2 == 0) ? 2 : 1); };

// 3-codon index macros (src/hmmer.h:312-314) with the reference's clamp to the degenerate rows (:347-349)
__device__ __forceinline__ int codon2_fs3(int w, int x)               { int c = x * 84 + w * 21;                 return min(c, 337); }
__device__ __forceinline__ int codon3_fs3(int v, int w, int x)        { int c = x * 84 + w * 21 + v * 5 + 1;     return min(c, 336); }
__device__ __forceinline__ int codon4_fs3(int u, int v, int w, int x) { int c = x * 84 + w * 21 + v * 5 + u + 2; return min(c, 337); }

template <int J, int VEC>
__device__ __forceinline__ void load_emission_row(const float *__restrict__ row, int lane, float (&e)[J])
{
#pragma unroll
  for (int g = 0; g < J / VEC; ++g) {
    if constexpr (VEC == 4) {
      float4 t = __ldg(reinterpret_cast<const float4 *>(row) + g * kWarp + lane);
      e[4 * g + 0] = t.x; e[4 * g + 1] = t.y; e[4 * g + 2] = t.z; e[4 * g + 3] = t.w;
    } else if constexpr (VEC == 2) {
      float2 t = __ldg(reinterpret_cast<const float2 *>(row) + g * kWarp + lane);
      e[2 * g + 0] = t.x; e[2 * g + 1] = t.y;
    } else {
      e[g] = __ldg(row + g * kWarp + lane);
    }
  }
}

__device__ __forceinline__ float warp_allsum(float v)
{
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// fetch the 16 bits (4 nucleotides) for rows i..i+3 of a window; p0 = 0-based block index of row i
__device__ __forceinline__ uint32_t fetch_quad(const uint32_t *__restrict__ dna4, long long p0)
{
  long long q  = p0 + 8;                 // guard word in front
  long long wi = q >> 3;
  int       sh = (int)(q & 7) * 4;
  uint32_t lo = __ldg(dna4 + wi), hi = __ldg(dna4 + wi + 1);
  return __funnelshift_r(lo, hi, sh) & 0xffffu;
}

template <int J>
struct FwdState {
  float V[4][J];       // V(r) in slot r&3
  float I[4][J];       // I(r) in slot r&3
  float xN[4], xJ[4], xC[4];
};

template <int J>
struct FwdConsts {
  float bm[J], mm[J], im[J], dm[J], md[J], dd[J], pp[J], mi[J], ii[J], z[J];
  float bs[5];
};

// One DP row i of the Forward parser.  PH = i & 3 (compile time).
// cend[0..2] capture C(L), C(L-1), C(L-2) because the row loop may run up to 3 rows past L.
template <int J, int VEC, int PH, bool XMX>
__device__ __forceinline__ void fwd_row(int i, int L, int lane, FwdState<J> &S, const FwdConsts<J> &K,
                                        const float *__restrict__ emis, int mpad, int c2, int c3, int c4,
                                        float ploop, float pmove, float tEL, float tEM,
                                        float &totscale, float (&cend)[3], float *__restrict__ xrow)
{
  constexpr int P0 = PH, P1 = (PH + 3) & 3, P2 = (PH + 2) & 3, P3 = (PH + 1) & 3;  // slots of rows i, i-1, i-2, i-3
  float e2[J], e3[J], e4[J], m[J];
  load_emission_row<J, VEC>(emis + (size_t)c2 * mpad, lane, e2);
  load_emission_row<J, VEC>(emis + (size_t)c3 * mpad, lane, e3);
  load_emission_row<J, VEC>(emis + (size_t)c4 * mpad, lane, e4);

  // M(i,k) = V(i) R2 + V(i-1) R3 + V(i-2) R4        (fwdback_fs.c:390-392)
  float es0 = 0.f, es1 = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    float t = S.V[P0][j] * e2[j];
    t = fmaf(S.V[P1][j], e3[j], t);
    t = fmaf(S.V[P2][j], e4[j], t);
    m[j] = t;
    if (j & 1) es1 = fmaf(t, K.z[j], es1); else es0 = fmaf(t, K.z[j], es0);
  }
  float xE = warp_allsum(es0 + es1);

  // D chain: lane-local pass, then warp scan of the lane carries   (:415-453)
  float dl[J];
  dl[0] = 0.f;
#pragma unroll
  for (int j = 0; j + 1 < J; ++j) dl[j + 1] = fmaf(dl[j], K.dd[j], m[j] * K.md[j]);
  float A = fmaf(dl[J - 1], K.dd[J - 1], m[J - 1] * K.md[J - 1]);
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    float up = __shfl_up_sync(0xffffffffu, A, 1 << s);
    A = fmaf(K.bs[s], up, A);
  }
  float xin = __shfl_up_sync(0xffffffffu, A, 1);
  if (lane == 0) xin = 0.f;

  // specials   (:462-465; rows 0..2 hold N at 1.0, :155,279)
  float xN = (i < 3) ? 1.0f : S.xN[P3] * ploop;
  float xJ = fmaf(S.xJ[P3], ploop, xE * tEL);
  float xC = fmaf(S.xC[P3], ploop, xE * tEM);
  float xB = fmaf(xJ, pmove, xN * pmove);

  // outgoing mass O(i)[k] = M tMM + I tIM + D tDM ; V(i+2)[k] = O(i)[k-1] + B(i) tBM[k-1]   (:383-387)
  // I(i+3)[k] = M(i,k) tMI + I(i,k) tII                                                       (:408-409)
  float o[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    float d = fmaf(xin, K.pp[j], dl[j]);
    float t = m[j] * K.mm[j];
    t = fmaf(S.I[P0][j], K.im[j], t);
    t = fmaf(d, K.dm[j], t);
    o[j] = t;
    S.I[P1][j] = fmaf(S.I[P0][j], K.ii[j], m[j] * K.mi[j]);     // slot (i+3)&3 == (i-1)&3
  }
  float oprev = __shfl_up_sync(0xffffffffu, o[J - 1], 1);
  if (lane == 0) oprev = 0.f;
  S.V[P2][0] = fmaf(xB, K.bm[0], oprev);
#pragma unroll
  for (int j = 1; j < J; ++j) S.V[P2][j] = fmaf(xB, K.bm[j], o[j - 1]);

  float scale = 1.0f;
  if (xE > 1.0e4f && i <= L) {   // sparse rescaling (:472-496); warp-uniform branch
    float sf = 1.0f / xE;
    scale = xE;
    xN *= sf; xJ *= sf; xC *= sf; xB *= sf;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < J; ++j) { S.V[r][j] *= sf; S.I[r][j] *= sf; }
      S.xN[r] *= sf; S.xJ[r] *= sf; S.xC[r] *= sf;
    }
    cend[1] *= sf; cend[2] *= sf;
    totscale += logf(xE);
    xE = 1.0f;
  }
  S.xN[P0] = xN; S.xJ[P0] = xJ; S.xC[P0] = xC;
  if (i == L)     cend[0] = xC;
  if (i == L - 1) cend[1] = xC;
  if (i == L - 2) cend[2] = xC;

  if constexpr (XMX) {
    if (lane == 0 && i <= L) {
      float2 *x2 = reinterpret_cast<float2 *>(xrow + (size_t)i * 6);
      x2[0] = make_float2(xE, xN);
      x2[1] = make_float2(xJ, xB);
      x2[2] = make_float2(xC, scale);
    }
  }
}

template <int J>
__device__ __forceinline__ void load_fwd_consts(const float *__restrict__ cc, int lane, FwdConsts<J> &K)
{
#pragma unroll
  for (int j = 0; j < J; ++j) {
    K.bm[j] = __ldg(cc + (FC_BM * J + j) * kWarp + lane);
    K.mm[j] = __ldg(cc + (FC_MM * J + j) * kWarp + lane);
    K.im[j] = __ldg(cc + (FC_IM * J + j) * kWarp + lane);
    K.dm[j] = __ldg(cc + (FC_DM * J + j) * kWarp + lane);
    K.md[j] = __ldg(cc + (FC_MD * J + j) * kWarp + lane);
    K.dd[j] = __ldg(cc + (FC_DD * J + j) * kWarp + lane);
    K.pp[j] = __ldg(cc + (FC_PP * J + j) * kWarp + lane);
    K.mi[j] = __ldg(cc + (FC_MI * J + j) * kWarp + lane);
    K.ii[j] = __ldg(cc + (FC_II * J + j) * kWarp + lane);
    K.z[j]  = __ldg(cc + (FC_Z  * J + j) * kWarp + lane);
  }
#pragma unroll
  for (int s = 0; s < 5; ++s) K.bs[s] = __ldg(cc + FC_COUNT * J * kWarp + (FL_B0 + s) * kWarp + lane);
}

template <int J, bool XMX>
__global__ void __launch_bounds__(128) fs3_forward_parser_kernel(FsParserArgs a)
{
  constexpr int VEC = VecOf<J>::V;
  const int lane = threadIdx.x & 31;

  FwdConsts<J> K;                      // per-lane profile constants -> registers, once per warp
  load_fwd_consts<J>(a.cellc, lane, K);

  for (;;) {
    int w = 0;
    if (lane == 0) w = atomicAdd(a.counter, 1);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= a.nwin) break;

    const WindowDesc wd = a.wins[w];
    const int   L     = wd.L;
    const float pmove = wd.pmove, ploop = wd.ploop;
    float *xrow = nullptr;
    if constexpr (XMX) xrow = a.xmx + (size_t)a.xoff[w] * 6;

    FwdState<J> S;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int j = 0; j < J; ++j) { S.V[r][j] = 0.f; S.I[r][j] = 0.f; }
      S.xN[r] = 0.f; S.xJ[r] = 0.f; S.xC[r] = 0.f;
    }
    float totscale = 0.f;
    float cend[3] = { 0.f, 0.f, 0.f };

    // rows 0..4*nq-1 >= L; rows past L run on whatever follows the window and are ignored
    const int       nq     = (L + 4) >> 2;
    const long long p_base = wd.start - 1;       // 0-based block index of window position 1
    int i = 0;
    int u = 338, v = 338, wn = 338;              // (n[i-3], n[i-2], n[i-1]); 338 = degenerate placeholder (:176)
    uint32_t quad = fetch_quad(a.dna4, p_base - 1);

#define BATHGPU_ROW(PH_)                                                                              \
    {                                                                                                 \
      int nt = (int)(quad & 15u); quad >>= 4;                                                         \
      int xn = (i >= 1 && nt < 4) ? nt : 338;                                                         \
      int c2 = codon2_fs3(wn, xn), c3 = codon3_fs3(v, wn, xn), c4 = codon4_fs3(u, v, wn, xn);         \
      fwd_row<J, VEC, PH_, XMX>(i, L, lane, S, K, a.emis, a.mpad, c2, c3, c4, ploop, pmove,           \
                                a.tEL, a.tEM, totscale, cend, xrow);                                  \
      u = v; v = wn; wn = xn; ++i;                                                                    \
    }
    for (int q = 0; q < nq; ++q) {
      uint32_t next = fetch_quad(a.dna4, p_base + (i + 3));
      BATHGPU_ROW(0) BATHGPU_ROW(1) BATHGPU_ROW(2) BATHGPU_ROW(3)
      quad = next;
    }
#undef BATHGPU_ROW

    // final score (:513-529): (C(L) + C(L-1) tCL + C(L-2) tCL) tCM
    {
      float tot = cend[0] + cend[1] * ploop + cend[2] * ploop;
      int   st  = 0;
      float sc;
      if (isnan(tot) || isinf(tot))  { st = 16; sc = tot; }
      else if (L > 2 && tot == 0.0f) { st = 16; sc = -INFINITY; }
      else sc = totscale + logf(tot * pmove);
      if (lane == 0) { a.fwdsc[w] = sc; a.status[w] = st; }
    }
  }
}

}  // namespace bathgpu
