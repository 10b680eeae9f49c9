// stotrace.cpp -- the multi-domain branch of frameshift domain definition, host side.
//
// When is_multidomain_region_frameshift fires (src/p7_domaindef.c:395), the reference fills a multihit Forward matrix for the
// region (:411-412), samples 200 tracebacks from it (region_trace_ensemble_frameshift, :892-954; p7_StochasticTrace_Frameshift,
// src/impl_sse/stotrace_fs.c:72-365), clusters the sampled domain coordinates (p7_spensemble_fs_Cluster,
// src/p7_spensemble.c:498-640) and rescores every significant cluster as its own envelope.  The matrix is the DP and comes
// from the device (bathgpu_fs_forward_matrices); the sampling is a serial random walk through it -- one dependent draw per
// step, the generator's state threading all 200 walks -- and stays here, as it stays host code in the reference's own layering
// (it is not part of the impl layer's kernel set), together with the integer clustering.
//
// A sampled trace is only ever reduced to its domains' end points (p7_trace_fs_Index, src/p7_trace.c:2645-2680), so the walk
// records those directly instead of building the trace.
//
// Easel pieces restated (Easel is not in the reference tree): esl_randomness_CreateFast / esl_random (32-bit LCG 69069 x + 1,
// Jenkins-mixed seed), esl_rnd_FChoose, esl_vec_FNorm (compensated sum), esl_cluster_SingleLinkage.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "host_internal.h"
#include "../../include/bathhost.h"

namespace bathhost {

namespace {

struct FastRng {
  uint32_t x;
  explicit FastRng(uint32_t seed)
  {
    uint32_t a = seed, b = 87654321u, c = 12345678u;
    a -= b; a -= c; a ^= (c >> 13);  b -= c; b -= a; b ^= (a << 8);   c -= a; c -= b; c ^= (b >> 13);
    a -= b; a -= c; a ^= (c >> 12);  b -= c; b -= a; b ^= (a << 16);  c -= a; c -= b; c ^= (b >> 5);
    a -= b; a -= c; a ^= (c >> 3);   b -= c; b -= a; b ^= (a << 10);  c -= a; c -= b; c ^= (b >> 15);
    x = c ? c : 42u;
  }
  double next() { x = x * 69069u + 1u; return (double) x / 4294967296.0; }
};

template <int N> int choose_normalised(FastRng &rng, float (&p)[N])
{
  float sum = 0.0f, comp = 0.0f;
  for (int z = 0; z < N; ++z) { const float y = p[z] - comp, t = sum + y; comp = (t - sum) - y; sum = t; }
  if (sum != 0.0f) for (float &v : p) v /= sum; else for (float &v : p) v = 1.0f / (float) N;
  const double roll = rng.next();
  double acc = 0.0;
  for (int z = 0; z < N; ++z) { acc += p[z]; if (roll < acc) return z; }
  int z;
  do { z = (int)(rng.next() * N); } while (p[z] == 0.0f);
  return z;
}

enum Cell { cD = 0, cI = 1, cM = 2 };                       // then M_C1..M_C5 at 3..7
enum XCell { xE = 0, xN, xJ, xB, xC, xS };
enum Tr { tBM = 0, tMM, tIM, tDM, tMD, tMI, tII, tDD };
enum St { sM, sD, sI, sS, sN, sB, sE, sC, sJ };

}  // namespace

// 200 (nsamples) tracebacks through the Forward matrix of region ireg..ireg+L-1; one Segment per sampled domain, in sampling
// order and, inside a trace, in sequence order.  false if a walk leaves the matrix (the reference would read out of bounds).
bool sample_region_segments(const ForwardMatrix &F, const float *tfv, const SpecialOdds &X, uint32_t seed, int nsamples, int ireg,
                            std::vector<Segment> &out)
{
  const int M = F.M, L = F.L, ld = M + 1;
  const int Q = std::max(2, (M - 1) / 4 + 1);               // p7O_NQF (impl_sse.h:26): select_e visits nodes in striped order
  auto cell = [&](int i, int k, int c) { return F.mx[((size_t) i * ld + k) * 8 + c]; };
  auto xr   = [&](int i, int c) { return F.xr[(size_t) i * 6 + c]; };
  auto T    = [&](int t, int k) { return tfv[(size_t) t * ld + k]; };
  FastRng rng(seed);
  out.clear();
  std::vector<Segment> doms;
  for (int t = 0; t < nsamples; ++t) {
    doms.clear();
    int i = L, k = 0, c = 0, s0 = sC, s1 = sC;
    Segment cur{ t, 0, 0, 0, 0, 0.0f };
    bool have_m = false;
    while (s0 != sS) {
      switch (s0) {
      case sM: {
        float p[4] = { xr(i, xB) * T(tBM, k - 1), 0.f, 0.f, 0.f };
        if (k > 1) { p[1] = cell(i, k - 1, cM) * T(tMM, k - 1); p[2] = cell(i, k - 1, cI) * T(tIM, k - 1); p[3] = cell(i, k - 1, cD) * T(tDM, k - 1); }
        static const int st[4] = { sB, sM, sI, sD };
        s1 = st[choose_normalised(rng, p)]; k--;
        break; }
      case sD: {
        float p[2] = { 0.f, 0.f };
        if (k > 1) { p[0] = cell(i, k - 1, cM) * T(tMD, k - 1); p[1] = cell(i, k - 1, cD) * T(tDD, k - 1); }
        s1 = choose_normalised(rng, p) == 0 ? sM : sD; k--;
        break; }
      case sI: {
        if (i < 3) return false;
        float p[2] = { cell(i - 3, k, cM) * T(tMI, k), cell(i - 3, k, cI) * T(tII, k) };
        s1 = choose_normalised(rng, p) == 0 ? sM : sI; i -= 3;
        break; }
      case sN: s1 = (i == 0) ? sS : sN; break;
      case sC: case sJ: {
        if (i < 4) { s1 = sE; break; }
        const int   xc = (s0 == sC) ? xC : xJ;
        const float e_odds = (s0 == sC) ? X.e_move : X.e_loop;
        const float s2 = xr(i - 2, xS), s1f = xr(i - 1, xS), s0f = xr(i, xS);
        float p[4] = { xr(i - 3, xc) * X.loop, xr(i - 2, xc) * X.loop * s2, xr(i - 1, xc) * X.loop * s2 * s1f, xr(i, xE) * e_odds * s2 * s1f * s0f };
        s1 = (choose_normalised(rng, p) < 3) ? s0 : sE;
        break; }
      case sE: {
        double sum = 0.0;
        const double roll = rng.next(), norm = 1.0 / xr(i, xE);
        const float nf = (float) norm;
        s1 = -1;
        for (int pass = 0; pass < 1000 && s1 < 0; ++pass)
          for (int q = 0; q < Q && s1 < 0; ++q) {
            for (int z = 0; z < 4 && s1 < 0; ++z) { const int kk = z * Q + q + 1; sum += (kk <= M) ? cell(i, kk, cM) * nf : 0.0f; if (roll < sum) { k = kk; s1 = sM; } }
            for (int z = 0; z < 4 && s1 < 0; ++z) { const int kk = z * Q + q + 1; sum += (kk <= M) ? cell(i, kk, cD) * nf : 0.0f; if (roll < sum) { k = kk; s1 = sD; } }
          }
        if (s1 < 0) return false;
        break; }
      case sB: {
        float p[2] = { xr(i, xN) * X.move, xr(i, xJ) * X.move };
        s1 = choose_normalised(rng, p) == 0 ? sN : sJ;
        break; }
      default: return false;
      }
      if (s1 == sM) {
        float p[5] = { cell(i, k, 3), cell(i, k, 4), cell(i, k, 5), cell(i, k, 6), cell(i, k, 7) };
        c = choose_normalised(rng, p) + 1;
        if (i - c < 0) s1 = sB;                             // codon would start before the region (stotrace_fs.c:111)
      } else c = 0;
      // what p7_trace_fs_Index keeps of the step just appended
      if (s1 == sE) { cur = Segment{ t, 0, 0, 0, 0, 0.0f }; have_m = false; }
      else if (s1 == sM) {
        if (!have_m) { cur.j = i; cur.m = k; have_m = true; }
        cur.i = i - c + 1; cur.k = k;
      } else if (s1 == sB) doms.push_back(cur);
      if ((s1 == sN || s1 == sC || s1 == sJ) && s1 == s0) i--;
      s0 = s1;
      i -= c;
      if (i < 0) return false;
    }
    for (size_t d = doms.size(); d-- > 0;) {
      Segment g = doms[d];
      g.i += ireg - 1; g.j += ireg - 1;
      out.push_back(g);
    }
  }
  return true;
}

// The standard-translation flavour: p7_StochasticTrace (src/impl_sse/stotrace.c:72-326) over the Forward matrix of a region of an
// ORF, nsamples times from one generator; every sampled trace is reduced to its domains (p7_trace_Index, src/p7_trace.c:2592-2625)
// and each domain's null2 odds are taken from the trace (p7_Null2_ByTrace, src/impl_sse/null2.c:131-219: emitting-state usage of the
// segment -- insert states are counted on their node's match cell, as the reference's workspace indexing does -- times the match
// emission odds, summed in the striped order of the SIMD code) and accumulated per residue as region_trace_ensemble does
// (src/p7_domaindef.c:785-815; the first residue of a domain is bumped by 1 like the residues outside domains, :797).
bool sample_region_segments_protein(const ForwardMatrix &F, const float *tfv, const float *rf, const SpecialOdds &X, uint32_t seed, int nsamples,
                                    int ireg, const uint8_t *res, std::vector<Segment> &out, std::vector<float> &n2sc)
{
  const int M = F.M, L = F.L, ld = M + 1;
  const int Q = std::max(2, (M - 1) / 4 + 1);               // p7O_NQF
  auto cell = [&](int i, int k, int c) { return F.mx[((size_t) i * ld + k) * 4 + c]; };      // c: 0 M, 1 D, 2 I
  auto xr   = [&](int i, int c) { return F.xr[(size_t) i * 6 + c]; };
  auto T    = [&](int t, int k) { return tfv[(size_t) t * ld + k]; };
  FastRng rng(seed);
  out.clear();
  n2sc.assign((size_t) L + 1, 0.0f);
  struct Dom { Segment g; float null2[kKp]; };
  std::vector<Dom> doms;
  std::vector<int> used;                                     // nodes of the emitting states of the domain being walked
  for (int t = 0; t < nsamples; ++t) {
    doms.clear();
    int i = L, k = 0, s0 = sC, s1 = sC;
    Segment cur{ t, 0, 0, 0, 0, 0.0f };
    bool have_m = false;
    used.clear();
    while (s0 != sS) {
      switch (s0) {
      case sM: {
        if (i < 1 || k < 1) return false;
        float p[4] = { xr(i - 1, xB) * T(tBM, k - 1), 0.f, 0.f, 0.f };
        if (k > 1) { p[1] = cell(i - 1, k - 1, 0) * T(tMM, k - 1); p[2] = cell(i - 1, k - 1, 2) * T(tIM, k - 1); p[3] = cell(i - 1, k - 1, 1) * T(tDM, k - 1); }
        static const int st[4] = { sB, sM, sI, sD };
        s1 = st[choose_normalised(rng, p)]; k--; i--;
        break; }
      case sD: {
        float p[2] = { 0.f, 0.f };
        if (k > 1) { p[0] = cell(i, k - 1, 0) * T(tMD, k - 1); p[1] = cell(i, k - 1, 1) * T(tDD, k - 1); }
        s1 = choose_normalised(rng, p) == 0 ? sM : sD; k--;
        break; }
      case sI: {
        if (i < 1) return false;
        float p[2] = { cell(i - 1, k, 0) * T(tMI, k), cell(i - 1, k, 2) * T(tII, k) };
        s1 = choose_normalised(rng, p) == 0 ? sM : sI; i--;
        break; }
      case sN: s1 = (i == 0) ? sS : sN; break;
      case sC: case sJ: {
        if (i < 1) { s1 = sE; break; }                       // C(0) = J(0) = 0: only E can have been the source (never reached with E(0) = 0)
        const int   xc = (s0 == sC) ? xC : xJ;
        const float e_odds = (s0 == sC) ? X.e_move : X.e_loop;
        float p[2] = { xr(i - 1, xc) * X.loop, xr(i, xE) * e_odds * xr(i, xS) };
        s1 = (choose_normalised(rng, p) == 0) ? s0 : sE;
        break; }
      case sE: {
        double sum = 0.0;
        const double roll = rng.next(), norm = 1.0 / xr(i, xE);
        const float nf = (float) norm;
        s1 = -1;
        for (int pass = 0; pass < 1000 && s1 < 0; ++pass)
          for (int q = 0; q < Q && s1 < 0; ++q) {
            for (int z = 0; z < 4 && s1 < 0; ++z) { const int kk = z * Q + q + 1; sum += (kk <= M) ? cell(i, kk, 0) * nf : 0.0f; if (roll < sum) { k = kk; s1 = sM; } }
            for (int z = 0; z < 4 && s1 < 0; ++z) { const int kk = z * Q + q + 1; sum += (kk <= M) ? cell(i, kk, 1) * nf : 0.0f; if (roll < sum) { k = kk; s1 = sD; } }
          }
        if (s1 < 0) return false;
        break; }
      case sB: {
        float p[2] = { xr(i, xN) * X.move, xr(i, xJ) * X.move };
        s1 = choose_normalised(rng, p) == 0 ? sN : sJ;
        break; }
      default: return false;
      }
      // what p7_trace_Index and p7_Null2_ByTrace keep of the step just appended (state s1 at node k, residue i)
      if (s1 == sE) { cur = Segment{ t, 0, 0, 0, 0, 0.0f }; have_m = false; used.clear(); }
      else if (s1 == sM) {
        if (!have_m) { cur.j = i; cur.m = k; have_m = true; }
        cur.i = i; cur.k = k;
        used.push_back(k);
      } else if (s1 == sI) used.push_back(k);
      else if (s1 == sB) {
        Dom d;
        d.g = cur;
        // p7_Null2_ByTrace over the domain's states
        std::vector<float> cnt((size_t) M + 1, 0.0f);
        for (int kk : used) cnt[(size_t) kk] += 1.0f;
        const float norm = 1.0f / (float) used.size();
        for (float &c : cnt) c *= norm;
        for (int x = 0; x < kK; ++x) {
          float part[4] = { 0.f, 0.f, 0.f, 0.f };
          for (int q = 0; q < Q; ++q)
            for (int z = 0; z < 4; ++z) { const int kk = z * Q + q + 1; if (kk <= M) part[z] += cnt[(size_t) kk] * rf[(size_t) x * ld + kk]; }
          d.null2[x] = (part[0] + part[1]) + (part[2] + part[3]);             // esl_sse_hsum_ps; N, C, J usage of a B..E segment is 0
        }
        // esl_abc_FAvgScVec: degenerate residues take the average over their members; gap, '*' and '~' are 1
        static const int members[6][2] = { { 2, 11 }, { 7, 9 }, { 3, 13 }, { 8, 8 }, { 1, 1 }, { -1, -1 } };
        for (int x = kK + 1; x <= kKp - 3; ++x) {
          const int *mb = members[x - kK - 1];
          float sum = 0.0f; int nmb = 0;
          for (int y = 0; y < kK; ++y) if (mb[0] < 0 || y == mb[0] || y == mb[1]) { sum += d.null2[y]; ++nmb; }
          d.null2[x] = sum / (float) nmb;
        }
        d.null2[kK] = 1.0f; d.null2[kKp - 2] = 1.0f; d.null2[kKp - 1] = 1.0f;
        doms.push_back(d);
      }
      if ((s1 == sN || s1 == sC || s1 == sJ) && s1 == s0) i--;
      s0 = s1;
      if (i < 0) return false;
    }
    // region_trace_ensemble's per-residue accumulation, domains in sequence order (the walk found them last to first)
    int pos = 1;
    for (size_t d = doms.size(); d-- > 0;) {
      const Dom &D = doms[d];
      for (; pos <= D.g.i; pos++) n2sc[(size_t) pos] += 1.0f;
      for (; pos <= D.g.j; pos++) n2sc[(size_t) pos] += D.null2[res[pos]];
      Segment g = D.g;
      g.i += ireg - 1; g.j += ireg - 1;
      out.push_back(g);
    }
    for (; pos <= L; pos++) n2sc[(size_t) pos] += 1.0f;
  }
  for (int pos = 1; pos <= L; ++pos) n2sc[(size_t) pos] = logf(n2sc[(size_t) pos] / (float) nsamples);
  return true;
}

// link_spsamples_fs (src/p7_spensemble.c:226-256): overlap >= 0.8 of the smaller segment on both axes and start or end
// within 4 diagonals
static bool linked(const Segment &a, const Segment &b, bool protein)
{
  const float min_overlap = 0.8f; const int max_diagdiff = 4;
  int nov = std::min(a.j, b.j) - std::max(a.i, b.i) + 1;
  int n   = std::min(a.j - a.i + 1, b.j - b.i + 1);
  if ((float) nov / (float) n < min_overlap) return false;
  nov = std::min(a.m, b.m) - std::max(a.k, b.k);
  n   = std::min(a.m - a.k + 1, b.m - b.k + 1);
  if ((float) nov / (float) n < min_overlap) return false;
  if (protein) {                                           // link_spsamples (src/p7_spensemble.c:191-218): residue coordinates
    if (std::abs((a.i - a.k) - (b.i - b.k)) <= max_diagdiff) return true;
    if (std::abs((a.j - a.m) - (b.j - b.m)) <= max_diagdiff) return true;
    return false;
  }
  if (std::abs((a.i / 3 - a.k) - (b.i / 3 - b.k)) <= max_diagdiff) return true;
  if (std::abs((a.j / 3 - a.m) - (b.j / 3 - b.m)) <= max_diagdiff) return true;
  return false;
}

// p7_spensemble_fs_Cluster with the parameters p7_domaindef_Create sets (src/p7_domaindef.c:83-88), followed by the removal of
// dominated clusters (:923-952).  Returns consensus segments ordered by start.
std::vector<Segment> cluster_region_segments(const std::vector<Segment> &sp, int nsamples, bool protein)
{
  const float min_posterior = 0.25f, min_endpointp = 0.02f;
  const int n = (int) sp.size();
  std::vector<int> pool(n), stack, asg(n, -1);
  for (int v = 0; v < n; ++v) pool[v] = n - v - 1;          // esl_cluster_SingleLinkage: vertex 0 is popped first
  int nc = 0;
  while (!pool.empty()) {
    stack.push_back(pool.back()); pool.pop_back();
    while (!stack.empty()) {
      const int v = stack.back(); stack.pop_back();
      asg[v] = nc;
      for (int z = (int) pool.size() - 1; z >= 0; --z)
        if (linked(sp[v], sp[pool[z]], protein)) { stack.push_back(pool[z]); pool[z] = pool.back(); pool.pop_back(); }
    }
    ++nc;
  }
  std::vector<Segment> sig;
  std::vector<int> epc;
  for (int c = 0; c < nc; ++c) {
    int ninc = 0, last = -1;
    for (int h = 0; h < n; ++h) if (asg[h] == c) { if (sp[h].idx != last) ++ninc; last = sp[h].idx; }
    if ((float) ninc / (float) nsamples < min_posterior) continue;
    int lo[4] = { 0, 0, 0, 0 }, hi[4] = { 0, 0, 0, 0 };
    bool first = true;
    auto coord = [](const Segment &g, int a) { return a == 0 ? g.i : a == 1 ? g.j : a == 2 ? g.k : g.m; };
    for (int h = 0; h < n; ++h) if (asg[h] == c)
      for (int a = 0; a < 4; ++a) {
        const int v = coord(sp[h], a);
        if (first) lo[a] = hi[a] = v; else { lo[a] = std::min(lo[a], v); hi[a] = std::max(hi[a], v); }
        if (a == 3) first = false;
      }
    const int thr = (int) ceilf((float) ninc * min_endpointp);
    int best[4];
    for (int a = 0; a < 4; ++a) {
      const int w = hi[a] - lo[a] + 1;
      epc.assign((size_t) w, 0);
      for (int h = 0; h < n; ++h) if (asg[h] == c) epc[coord(sp[h], a) - lo[a]]++;
      const bool leftmost = (a == 0 || a == 2);              // i and k: widest = leftmost; j and m: rightmost
      int pick = -1;
      if (leftmost) { for (int z = 0; z < w; ++z) if (epc[z] >= thr) { pick = z; break; } }
      else          { for (int z = w - 1; z >= 0; --z) if (epc[z] >= thr) { pick = z; break; } }
      if (pick < 0) pick = (int)(std::max_element(epc.begin(), epc.end()) - epc.begin());
      best[a] = lo[a] + pick;
    }
    if (best[0] > best[1] || best[2] > best[3]) continue;
    sig.push_back(Segment{ c, best[0], best[1], best[2], best[3], (float) ninc / (float) nsamples });
  }
  // qsort by start in the reference; ties keep cluster order here
  std::stable_sort(sig.begin(), sig.end(), [](const Segment &a, const Segment &b) { return a.i < b.i; });
  std::vector<char> dominated(sig.size(), 0);
  for (size_t d = 0; d < sig.size(); ++d)
    for (size_t d2 = d + 1; d2 < sig.size(); ++d2) {
      const int nov = std::min(sig[d].j, sig[d2].j) - std::max(sig[d].i, sig[d2].i) + 1;
      if (nov == 0) break;
      const int nn = std::min(sig[d].j - sig[d].i + 1, sig[d2].j - sig[d2].i + 1);
      if ((float) nov / (float) nn >= 0.8f) { if (sig[d].prob > sig[d2].prob) dominated[d2] = 1; else dominated[d] = 1; }
    }
  std::vector<Segment> out;
  for (size_t d = 0; d < sig.size(); ++d) if (!dominated[d]) out.push_back(sig[d]);
  return out;
}

}  // namespace bathhost

// ---- C entry points (include/bathhost.h): the two steps on caller-provided arrays, for tests and for a reference-side caller
extern "C" int bathhost_sample_region_segments(const float *mx, const float *xrows, int M, int L, const float *tfv, const float odds[4],
                                               uint32_t seed, int nsamples, int ireg, bathhost_segment *out, int max_out, int *nout)
{
  if (!mx || !xrows || !tfv || !odds || !out || !nout || M < 1 || L < 1 || nsamples < 1) return BATHHOST_EINVAL;
  std::vector<bathhost::Segment> sp;
  const bathhost::ForwardMatrix F{ mx, xrows, M, L };
  const bathhost::SpecialOdds X{ odds[0], odds[1], odds[2], odds[3] };
  if (!bathhost::sample_region_segments(F, tfv, X, seed, nsamples, ireg, sp)) return BATHHOST_EINVAL;
  *nout = (int) sp.size();
  if ((int) sp.size() > max_out) return BATHHOST_EINVAL;
  for (size_t z = 0; z < sp.size(); ++z) out[z] = bathhost_segment{ sp[z].idx, sp[z].i, sp[z].j, sp[z].k, sp[z].m, sp[z].prob };
  return BATHHOST_OK;
}

extern "C" int bathhost_cluster_region_segments(const bathhost_segment *sp, int n, int nsamples, bathhost_segment *out, int max_out, int *nout)
{
  if ((!sp && n > 0) || !out || !nout || n < 0 || nsamples < 1) return BATHHOST_EINVAL;
  std::vector<bathhost::Segment> in((size_t) n);
  for (int z = 0; z < n; ++z) in[z] = bathhost::Segment{ sp[z].idx, sp[z].i, sp[z].j, sp[z].k, sp[z].m, sp[z].prob };
  const std::vector<bathhost::Segment> sig = bathhost::cluster_region_segments(in, nsamples);
  *nout = (int) sig.size();
  if ((int) sig.size() > max_out) return BATHHOST_EINVAL;
  for (size_t z = 0; z < sig.size(); ++z) out[z] = bathhost_segment{ sig[z].idx, sig[z].i, sig[z].j, sig[z].k, sig[z].m, sig[z].prob };
  return BATHHOST_OK;
}

extern "C" int bathhost_sample_region_segments_protein(const float *mx, const float *xrows, int M, int L, const float *tfv, const float *rf,
                                                       const float odds[4], uint32_t seed, int nsamples, int ireg, const uint8_t *res,
                                                       bathhost_segment *out, int max_out, int *nout, float *n2sc)
{
  if (!mx || !xrows || !tfv || !rf || !odds || !res || !out || !nout || !n2sc || M < 1 || L < 1 || nsamples < 1) return BATHHOST_EINVAL;
  std::vector<bathhost::Segment> sp;
  std::vector<float> n2;
  const bathhost::ForwardMatrix F{ mx, xrows, M, L };
  const bathhost::SpecialOdds X{ odds[0], odds[1], odds[2], odds[3] };
  if (!bathhost::sample_region_segments_protein(F, tfv, rf, X, seed, nsamples, ireg, res, sp, n2)) return BATHHOST_EINVAL;
  *nout = (int) sp.size();
  if ((int) sp.size() > max_out) return BATHHOST_EINVAL;
  for (size_t z = 0; z < sp.size(); ++z) out[z] = bathhost_segment{ sp[z].idx, sp[z].i, sp[z].j, sp[z].k, sp[z].m, sp[z].prob };
  for (int p = 0; p <= L; ++p) n2sc[p] = n2[(size_t) p];
  return BATHHOST_OK;
}

extern "C" int bathhost_cluster_region_segments_protein(const bathhost_segment *sp, int n, int nsamples, bathhost_segment *out, int max_out, int *nout)
{
  if ((!sp && n > 0) || !out || !nout || n < 0 || nsamples < 1) return BATHHOST_EINVAL;
  std::vector<bathhost::Segment> in((size_t) n);
  for (int z = 0; z < n; ++z) in[z] = bathhost::Segment{ sp[z].idx, sp[z].i, sp[z].j, sp[z].k, sp[z].m, sp[z].prob };
  const std::vector<bathhost::Segment> sig = bathhost::cluster_region_segments(in, nsamples, true);
  *nout = (int) sig.size();
  if ((int) sig.size() > max_out) return BATHHOST_EINVAL;
  for (size_t z = 0; z < sig.size(); ++z) out[z] = bathhost_segment{ sig[z].idx, sig[z].i, sig[z].j, sig[z].k, sig[z].m, sig[z].prob };
  return BATHHOST_OK;
}
