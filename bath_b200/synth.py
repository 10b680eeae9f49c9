"""Synthetic targets of the shapes BASELINE.json names: iid-ACGT genomes with planted,
back-translated, frameshifted homologs of a query model.  Seeded; no files, no network.

The recipe mirrors how the reference's own unit tests make homologous DNA: emit residues from the
model, back-translate with uniformly chosen synonymous codons (p7_codontable_GetCodon,
src/hmmer.c:258), then damage codons at the model's frameshift rate.
"""
import numpy as np

STD_CODE = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"   # codon 16*n1+4*n2+n3, ACGT
AMINO = "ACDEFGHIKLMNPQRSTVWY"


def _codon_table():
    """[20][6] codon indices per amino acid (padded by repetition) and [20] counts"""
    tab = np.zeros((20, 6), np.int64)
    cnt = np.zeros(20, np.int64)
    for a, sym in enumerate(AMINO):
        cs = [i for i, s in enumerate(STD_CODE) if s == sym]
        cnt[a] = len(cs)
        tab[a] = (cs * 6)[:6]
    return tab, cnt


def iid_genome(rng, n):
    """ESL_DSQ-style digital sequence: codes 0..3 at [1..n], sentinel 255 at [0] and [n+1]."""
    d = np.empty(n + 2, np.uint8)
    d[0] = d[n + 1] = 255
    d[1:-1] = rng.integers(0, 4, n, dtype=np.uint8)
    return d


def homolog(rng, mat, fs_rate=0.01, stop_every=300):
    """One DNA homolog of a model with match emissions mat[1..M][20]: a residue per match state,
    a random synonymous codon each, +-1/+-2 nt frameshifts at fs_rate per codon, a stop codon every
    ~stop_every codons."""
    tab, cnt = _codon_table()
    M = mat.shape[0] - 1
    cdf = np.cumsum(mat[1:].astype(np.float64), axis=1)
    cdf /= cdf[:, -1:]
    res = (rng.random((M, 1)) > cdf).sum(axis=1).clip(0, 19)
    codon = tab[res, (rng.random(M) * cnt[res]).astype(np.int64)]
    nts = np.stack([codon // 16, (codon // 4) % 4, codon % 4], axis=1).astype(np.uint8)
    stop = rng.random(M) < 1.0 / stop_every
    nts[stop] = (3, 0, 0)       # TAA
    keep = np.full(M, 3)        # nucleotides kept of each codon; >3 = random insertions after it
    hit = rng.random(M) < fs_rate
    keep[hit] = rng.choice([1, 2, 4, 5], size=int(hit.sum()))
    # assemble: undamaged stretches are copied whole
    pieces = []
    last = 0
    for k in np.nonzero(hit)[0]:
        pieces.append(nts[last:k].reshape(-1))
        c = nts[k]
        if keep[k] < 3:
            pieces.append(c[: keep[k]])
        else:
            pieces.append(np.concatenate([c, rng.integers(0, 4, keep[k] - 3, dtype=np.uint8)]))
        last = k + 1
    pieces.append(nts[last:].reshape(-1))
    return np.concatenate(pieces).astype(np.uint8)


def planted_genome(rng, n, mat, every=50000, fs_rate=0.01, stop_every=300, revcomp_fraction=0.5):
    """iid-ACGT genome of n nt with one homolog planted every `every` nt (half of them on the
    bottom strand).  Returns (dsq, plants) with plants = [(start, end, strand)] in 1-based coordinates."""
    d = iid_genome(rng, n)
    plants = []
    pos = every // 2
    comp = np.array([3, 2, 1, 0], np.uint8)
    while True:
        h = homolog(rng, mat, fs_rate, stop_every)
        if pos + len(h) + 1 > n:
            break
        strand = 1
        if rng.random() < revcomp_fraction:
            h = comp[h[::-1]]
            strand = -1
        d[pos: pos + len(h)] = h
        plants.append((pos, pos + len(h) - 1, strand))
        pos += every
    return d, plants


def planted_contigs(rng, total, mats, every=50000, fs_rates=0.01, min_len=1_000_000, max_len=10_000_000):
    """BASELINE config 4's target: iid-ACGT contigs of min_len..max_len nt adding up to `total`, a homolog planted every `every` nt,
    taken in turn from the models whose match emissions are in `mats` (half on the bottom strand).  Returns
    [(name, dsq)], [(contig, start, end, strand, model)]."""
    if not isinstance(fs_rates, (list, tuple)):
        fs_rates = [fs_rates] * len(mats)
    comp = np.array([3, 2, 1, 0], np.uint8)
    contigs, plants, left, turn = [], [], int(total), 0
    while left > 0:
        n = int(min(left, rng.integers(min_len, max_len + 1)))
        if left - n < min_len // 2:
            n = left
        d = iid_genome(rng, n)
        pos = every // 2
        while True:
            k = turn % len(mats)
            h = homolog(rng, mats[k], fs_rates[k])
            if pos + len(h) + 1 > n:
                break
            strand = 1
            if rng.random() < 0.5:
                h = comp[h[::-1]]
                strand = -1
            d[pos: pos + len(h)] = h
            plants.append((len(contigs), pos, pos + len(h) - 1, strand, k))
            pos += every
            turn += 1
        contigs.append((f"contig{len(contigs) + 1}", d))
        left -= n
    return contigs, plants


def tile_windows(n, length, step=None):
    """Window starts/lengths tiling [1..n]: the last window is pulled back so that every window is full length."""
    step = step or length
    starts = np.arange(1, max(n - length + 1, 1) + 1, step, dtype=np.int64)
    if starts[-1] + length - 1 < n:
        starts = np.append(starts, n - length + 1)
    return starts, np.full(len(starts), min(length, n), np.int32)
