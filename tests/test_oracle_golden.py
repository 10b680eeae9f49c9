"""Pins the CPU oracle on the reference's own golden output (tutorial/AMP_N-fs.out|.tbl).

The reference cannot be compiled here (Easel is not vendored), so the oracle is checked against what the
reference binary printed for `bathsearch --fs --cigar AMP_N.bhmm target-AMP_N.fa`: the hit's model and
target coordinates, the posterior-probability line, the translated-residue line, frameshift and stop
counts, percent identity and the CIGAR string.  Those are functions of: the .bhmm reader, the 5-codon
frameshift profile (scores AND the codons[]/indel_pos[] tables), Forward, Backward, Decoding, the
optimal-accuracy fill and its traceback -- i.e. every oracle function the GPU stages are compared with.
The alignment-display logic below restates p7_alidisplay_fs_Create (src/p7_alidisplay.c:696-823).
"""
import ctypes as C
import re

import numpy as np

import common

P__X, PX__, PXX_, PX_X, P_XX, PXXX, PXXx, PXxX, PxXX, Pxxx, PXXxX, PXxXX, PxXXX, PXXxxX, PXxxXX, PxxXXX = range(16)
AMINO = "ACDEFGHIKLMNPQRSTVWY-BJZOUX*~"


def codon_index5(nts):
    """get_codon_index (src/p7_alidisplay.c:32-88) with the fs5 row numbering (src/hmmer.h:306-310)"""
    c = len(nts)
    if any(n >= 4 for n in nts):
        return {1: 1366, 2: 1365, 3: 1364, 4: 1365, 5: 1366}[c]
    if c == 1:
        return nts[0] * 341
    if c == 2:
        w, x = nts
        return x * 341 + w * 85 + 1
    if c == 3:
        v, w, x = nts
        return x * 341 + w * 85 + v * 21 + 2
    if c == 4:
        u, v, w, x = nts
        return x * 341 + w * 85 + v * 21 + u * 5 + 3
    t, u, v, w, x = nts
    return x * 341 + w * 85 + v * 21 + u * 5 + t + 4


def encode_pp(p):
    """p7_alidisplay_EncodePostProb (src/p7_alidisplay.c:3689-3692)"""
    return "*" if p + 0.05 >= 1.0 else chr(int((p + 0.05) * 10.0) + ord("0"))


def display(trace, dsq, codons, indel_pos, consensus):
    """-> dict(aseq, ppline, cigar, shifts, stops, pid, hmm/ali coordinates) for the M/D/I part of a trace"""
    core = [t for t in trace if t[0] in "MDI"]
    aseq, pp, cigar = [], [], []
    shifts = stops = exact = 0
    n_count = 0
    for z, (st, k, i, c, p) in enumerate(core):
        nxt = core[z + 1][0] if z + 1 < len(core) else "E"
        pp.append("." if st == "D" else encode_pp(p))
        if st == "M":
            nts = [int(dsq[i - c + 1 + j]) for j in range(c)]
            ci = codon_index5(nts)
            aa, indel = int(codons[k, ci]), int(indel_pos[k, ci])
            aseq.append(AMINO[aa])
            if AMINO[aa].lower() == consensus[k].lower():
                exact += 1
            if c != 3:
                shifts += 1
            elif indel in (PXXx, PXxX, PxXX):
                stops += 1
            if nxt != "M" or c != 3:
                if c == 3:
                    n_count += 3
                elif indel in (PXX_, PXXxX, PXXxxX):
                    n_count += 2
                elif indel in (PX_X, PX__, PXxXX, PXxxXX):
                    n_count += 1
                cigar.append(f"{n_count}M")
                n_count = 0
                if c != 3:
                    cigar.append({1: "2B", 2: "1B", 4: "1F", 5: "2F"}[c])
                if indel in (P__X, PX_X, PXXxX, PXXxxX):
                    n_count = 1
                if indel in (P_XX, PXxXX, PXxxXX):
                    n_count = 2
                if indel in (PxXXX, PxxXXX):
                    n_count = 3
                if nxt != "M" and n_count > 0:
                    cigar.append(f"{n_count}M")
                    n_count = 0
            else:
                n_count += 3
        elif st == "I":
            ci = codon_index5([int(dsq[i - 2]), int(dsq[i - 1]), int(dsq[i])])
            indel = int(indel_pos[k, ci])
            if indel in (PXXx, PXxX, PxXX):
                stops += 1
                aseq.append("*")
            else:
                aseq.append(AMINO[int(codons[k, ci])].lower())
            n_count += 3
            if nxt != "I":
                cigar.append(f"{n_count}I")
                n_count = 0
        else:
            aseq.append("-")
            n_count += 3
            if nxt != "D":
                cigar.append(f"{n_count}D")
                n_count = 0
    first_m = next(t for t in core if t[0] == "M")
    emit = [t for t in core if t[0] != "D"]
    return {"aseq": "".join(aseq), "ppline": "".join(pp), "cigar": "".join(cigar), "shifts": shifts, "stops": stops,
            "pid": 100.0 * exact / len(core), "hmm_from": core[0][1], "hmm_to": core[-1][1],
            "ali_from": first_m[2] - first_m[3] + 1 if core[0][0] == "M" else emit[0][2] - 2, "ali_to": emit[-1][2]}


def golden_alignment(path):
    """aseq and PP lines of the single hit in AMP_N-fs.out (blocks of: model / match / aseq / ntseq / PP)"""
    lines = open(path).read().split("\n")
    aseq, pp = [], []
    for n, line in enumerate(lines):
        if re.match(r"^\s+seq1\s+\d+\s", line):
            aseq += lines[n - 1].split()
            pp += lines[n + 1].split()[:-1]
    return "".join(aseq), "".join(pp)


def domain_stage(po, model, dsq, n):
    L = po.lib()
    L.bo_fs_oprofile_ReconfigUnihit(model.om_fs5, n // 3)
    fwd, bck, oa = L.bo_mx_create(model.M, n, 8), L.bo_mx_create(model.M, n, 3), L.bo_mx_create(model.M, n, 3)
    fsc, bsc, e = C.c_float(), C.c_float(), C.c_float()
    assert L.bo_Forward_Frameshift(po.u8ptr(dsq), n, model.om_fs5, fwd, C.byref(fsc)) == 0
    assert L.bo_Backward_Frameshift(po.u8ptr(dsq), n, model.om_fs5, fwd, bck, C.byref(bsc)) == 0
    assert L.bo_Decoding_Frameshift(model.om_fs5, fwd, bck) == 0
    assert L.bo_OptimalAccuracy_Frameshift(model.om_fs5, fwd, oa, C.byref(e)) == 0
    tr = L.bo_trace_create()
    assert L.bo_OATrace_Frameshift(model.om_fs5, fwd, oa, tr) == 0
    trace = po.trace_list(tr)
    for mx in (fwd, bck, oa):
        L.bo_mx_destroy(mx)
    L.bo_trace_destroy(tr)
    L.bo_fs_oprofile_ReconfigMultihit(model.om_fs5, 100)
    return fsc.value, bsc.value, e.value, trace


def test_amp_n_alignment_matches_reference_output(oracle):
    po = oracle
    model = po.Model(common.golden("AMP_N.bhmm"))
    _, _, seq = po.read_fasta(common.golden("target-AMP_N.fa"))[0]
    dsq = po.digitize_dna(seq)
    fsc, bsc, oasc, trace = domain_stage(po, model, dsq, len(seq))
    assert abs(fsc - bsc) < 1e-3            # the reference's own fwd/bck bar (generic_fwdback_frameshift.c:2304-2435)

    gm = model.gm_fs5.contents
    mc = gm.maxcodons
    codons = np.ctypeslib.as_array(gm.codons, shape=((model.M + 1) * (mc + 1),))[: (model.M + 1) * mc].reshape(model.M + 1, mc)
    indel = np.ctypeslib.as_array(gm.indel_pos, shape=((model.M + 1) * (mc + 1),))[: (model.M + 1) * mc].reshape(model.M + 1, mc)
    consensus = model.hmm.contents.consensus.decode()
    d = display(trace, dsq, codons, indel, consensus)

    tbl = [l for l in open(common.golden("AMP_N-fs.tbl")) if not l.startswith("#")][0].split()
    # hit ID, target, acc, query, acc, hmm len, hmm from, hmm to, seq len, ali from, ali to, E, score, bias, PID, shifts, stops, CIGAR
    assert (d["hmm_from"], d["hmm_to"]) == (int(tbl[6]), int(tbl[7]))
    assert (d["ali_from"], d["ali_to"]) == (int(tbl[9]), int(tbl[10]))
    assert d["shifts"] == int(tbl[15]) and d["stops"] == int(tbl[16])
    assert d["cigar"] == tbl[17]
    assert f"{d['pid']:.2f}" == tbl[14]

    g_aseq, g_pp = golden_alignment(common.golden("AMP_N-fs.out"))
    assert d["aseq"] == g_aseq
    assert d["ppline"] == g_pp


def test_parser_and_full_matrix_agree(oracle):
    """parser == full matrix and fwd == bck, the invariants the reference's utests assert
    (src/impl_sse/fwdback_fs.c:3191-3258), on a frameshifted homolog and on random DNA."""
    po = oracle
    L = po.lib()
    model = po.Model(common.golden("2OG-FeII_Oxy_3.bhmm"))
    rng = np.random.default_rng(3)
    hom = common.sample_homolog(rng, common.hmm_mat(model), fs_rate=0.03)
    for dsq in (common.embed(rng, hom, 40, 51), common.random_dna(rng, 300)):
        n = len(dsq) - 2
        L.bo_fs_oprofile_ReconfigLength(model.om_fs3, n // 3)
        oxf, oxb = L.bo_mx_create(model.M, n, 0), L.bo_mx_create(model.M, n, 0)
        f3, b3 = C.c_float(), C.c_float()
        assert L.bo_ForwardParser_Frameshift_3Codons(po.u8ptr(dsq), n, model.om_fs3, oxf, C.byref(f3)) == 0
        assert L.bo_BackwardParser_Frameshift_3Codons(po.u8ptr(dsq), n, model.om_fs3, oxf, oxb, C.byref(b3)) == 0
        assert abs(f3.value - b3.value) < 1e-3
        L.bo_fs_oprofile_ReconfigLength(model.om_fs5, n // 3)
        fwd, bck = L.bo_mx_create(model.M, n, 8), L.bo_mx_create(model.M, n, 3)
        f5, b5 = C.c_float(), C.c_float()
        assert L.bo_Forward_Frameshift(po.u8ptr(dsq), n, model.om_fs5, fwd, C.byref(f5)) == 0
        assert L.bo_Backward_Frameshift(po.u8ptr(dsq), n, model.om_fs5, fwd, bck, C.byref(b5)) == 0
        assert abs(f5.value - b5.value) < 1e-3
        assert L.bo_Decoding_Frameshift(model.om_fs5, fwd, bck) == 0
        pp = po.mx_dp(fwd)
        # posterior rows: every nucleotide is emitted by exactly one codon/insert/flank per frame; PP cells are probabilities
        assert np.all(pp[1:, 1:, :] >= -1e-5) and np.all(pp[1:, 1:, :] <= 1.0 + 1e-3)
        for mx in (oxf, oxb, fwd, bck):
            L.bo_mx_destroy(mx)


def test_logsum_table(oracle):
    """p7_FLogsum (src/logsum.c:104-111): table-driven log(e^a + e^b), 0.001-nat bins"""
    L = oracle.lib()
    for a, b in [(0.0, 0.0), (-1.0, -3.5), (2.0, -20.0), (-5.0, -5.001)]:
        assert abs(L.bo_FLogsum(a, b) - np.logaddexp(a, b)) < 2e-3
    assert L.bo_FLogsum(-np.inf, -2.0) == -2.0
