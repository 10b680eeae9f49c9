"""Shared helpers for the tests: fixture paths and seeded synthetic inputs.

Nothing here reads /root/reference: the .bhmm / .fa fixtures the tests need are committed
under tests/golden/ (copied there by tests/golden/make_fixtures.py, which runs in the
build container only).
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

STD_CODE = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"
AA = "ACDEFGHIKLMNPQRSTVWY"


SYNTHETIC = {   # long models for the kernels' large-M instantiations: node blocks of shipped models concatenated
    "synthetic_M377.bhmm": [("tRNA-synthetases.bhmm", 0), ("tRNA-synthetases.bhmm", 1)],
    "synthetic_M208.bhmm": [("tRNA-proteins.bhmm", 11), ("tRNA-proteins.bhmm", 1)],
    "synthetic_M624.bhmm": [("tRNA-synthetases.bhmm", 0), ("tRNA-synthetases.bhmm", 1), ("tRNA-synthetases.bhmm", 2)],
    "synthetic_M903.bhmm": [("tRNA-synthetases.bhmm", 0), ("tRNA-synthetases.bhmm", 1), ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0)],
}


def _records(path):
    out, cur = [], []
    for line in open(path):
        cur.append(line)
        if line.startswith("//"):
            out.append(cur)
            cur = []
    return out


def concat_models(parts, out_path, name):
    """One profile whose nodes are those of the given (file, index) models in a row.  The last node of every part but the final
    one takes the transition line of the node before it (a model's last node has no delete / insert continuation).  Header,
    statistics and the node-0 lines come from the first part."""
    recs = [_records(golden(f))[i] for f, i in parts]
    head_end = next(z for z, l in enumerate(recs[0]) if l.startswith("HMM "))
    header = [l for l in recs[0][:head_end] if not l.startswith(("LENG", "MAXL", "NAME"))]
    nodes = []
    for r, rec in enumerate(recs):
        h = next(z for z, l in enumerate(rec) if l.startswith("HMM "))
        body = rec[h + 2:-1]
        if body[0].split()[0] == "COMPO":
            first = body[:3]
            body = body[3:]
        else:
            first = body[:2]
            body = body[2:]
        if r == 0:
            node0 = first
        blocks = [body[z:z + 3] for z in range(0, len(body), 3)]
        if r + 1 < len(recs):
            blocks[-1] = [blocks[-1][0], blocks[-1][1], blocks[-2][2]]
        nodes += blocks
    with open(out_path, "w") as f:
        f.write(header[0])
        f.write(f"NAME  {name}\nLENG  {len(nodes)}\n")
        f.writelines(header[1:])
        f.writelines(recs[0][head_end:head_end + 2])
        f.writelines(node0)
        for k, blk in enumerate(nodes, 1):
            fields = blk[0].split()
            f.write(f"{k:7d}   " + "  ".join(fields[1:21]) + "      - " + fields[22] + " - \n" if len(fields) >= 23 else blk[0])
            f.write(blk[1])
            f.write(blk[2])
        f.write("//\n")


def golden(name):
    path = os.path.join(GOLDEN, name)
    if name in SYNTHETIC and not os.path.exists(path):
        concat_models(SYNTHETIC[name], path, name.split(".")[0])
    return path


def random_dna(rng, n, p_degenerate=0.0):
    """iid ACGT (the reference's own benchmark distribution, fwdback_fs.c:3095), as ESL_DSQ."""
    d = np.full(n + 2, 255, np.uint8)
    d[1:-1] = rng.integers(0, 4, n)
    if p_degenerate > 0:
        mask = rng.random(n) < p_degenerate
        d[1:-1][mask] = 15  # N
    return d


def sample_homolog(rng, mat, fs_rate=0.01, stop_rate=0.003, sub_from_model=True):
    """A DNA sequence homologous to a profile: sample one residue per match state from
    mat[k] (k=1..M), back-translate with uniformly chosen synonymous codons (as
    p7_codontable_GetCodon, src/hmmer.c:258), then inject +-1/+-2 nt frameshifts at
    fs_rate per codon and occasional stop codons."""
    codons_for = {a: [] for a in AA}
    for idx, a in enumerate(STD_CODE):
        if a != "*":
            codons_for[a].append(idx)
    M = mat.shape[0] - 1
    out = []
    for k in range(1, M + 1):
        p = mat[k].astype(np.float64)
        p /= p.sum()
        a = AA[rng.choice(20, p=p)] if sub_from_model else AA[int(np.argmax(p))]
        c = codons_for[a][rng.integers(len(codons_for[a]))]
        nts = [c // 16, (c // 4) % 4, c % 4]
        if rng.random() < stop_rate:
            nts = [3, 0, 0]  # TAA
        r = rng.random()
        if r < fs_rate:
            kind = rng.integers(4)
            if kind == 0:
                nts = nts[:2]
            elif kind == 1:
                nts = nts[:1]
            elif kind == 2:
                nts = nts + [int(rng.integers(4))]
            else:
                nts = nts + [int(rng.integers(4)), int(rng.integers(4))]
        out += nts
    return np.array(out, np.uint8)


def embed(rng, insert, flank_left, flank_right):
    """random flank + insert + random flank, as ESL_DSQ (sentinels at both ends)"""
    n = flank_left + len(insert) + flank_right
    d = np.full(n + 2, 255, np.uint8)
    d[1:1 + flank_left] = rng.integers(0, 4, flank_left)
    d[1 + flank_left:1 + flank_left + len(insert)] = insert
    d[1 + flank_left + len(insert):-1] = rng.integers(0, 4, flank_right)
    return d


def hmm_mat(model):
    """match emission probabilities [M+1][20] of an oracle Model"""
    h = model.hmm.contents
    return np.ctypeslib.as_array(h.mat, shape=(h.M + 1, 20)).copy()
