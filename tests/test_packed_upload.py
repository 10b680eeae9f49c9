"""Host-packed target blocks (two nucleotides per byte): the form north_star names at the boundary ("packed 2-bit/4-bit DNA windows").
bathgpu_pack_dna4 on the host (CPU test), then on the device: a block uploaded packed must behave exactly like the same block
uploaded as ESL_DSQ bytes -- Forward scores of the one-call path bit for bit, the byte form the ORF finder reads, the reverse
complement made from it."""
import ctypes as C

import numpy as np
import pytest

import common


def test_pack_dna4_layout():
    """dsq[2j+1] in the low nibble of byte j, dsq[2j+2] in the high one, codes above 15 -> 15, odd tail nibble 15"""
    from bath_b200 import capi
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 8, 9, 1001):
        body = rng.integers(0, 18, n).astype(np.uint8)
        dsq = np.concatenate([[255], body, [255]]).astype(np.uint8)
        got = capi.pack_dna4(dsq)
        assert len(got) == (n + 1) // 2 == capi.load().bathgpu_packed4_bytes(n)
        c = np.minimum(body, 15)
        if n & 1:
            c = np.concatenate([c, [15]])
        assert np.array_equal(got, (c[0::2] | (c[1::2] << 4)).astype(np.uint8))
    L = capi.load()
    assert L.bathgpu_pack_dna4(None, 5, None) == 11 and L.bathgpu_packed4_bytes(-3) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("n", [9_000_001, 123_457, 64])
def test_packed_block_scores_equal_byte_block(oracle, gpu_ctx, n):
    """bathgpu_fs_fwd_block_packed4 == bathgpu_fs_fwd_block == bathgpu_upload_block_packed4 + bathgpu_fs_fwd_windows, bit for bit
    (odd lengths, three upload chunks at 9 M nucleotides, windows on the chunk borders and at both ends)"""
    from bath_b200 import capi
    model = oracle.Model(common.golden("AMP_N.bhmm"))
    rng = np.random.default_rng(n)
    dsq = common.random_dna(rng, n, p_degenerate=0.001)
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    if n > 2000:
        starts = np.sort(rng.integers(1, n - 1500, 3000))
        lens = rng.integers(60, 1300, len(starts))
        starts[:4] = [1, 2, n - 1299, n - 1199]; lens[:4] = [900, 61, 1300, 1200]
        if n > 6_000_000:
            starts[4:10] = [1_124_500, 1_124_990, 1_125_000, 5_625_000 - 700, 1_048_570, 5_242_880 - 700]
    else:
        starts, lens = np.array([1, 2, 5]), np.array([64, 63, 60])
    perm = rng.permutation(len(starts))
    w = capi.Context.make_windows(starts[perm], lens[perm])
    sc0, st0 = np.empty(len(w), np.float32), np.empty(len(w), np.int32)
    gpu_ctx.fs_fwd_block_into(dsq, w, (0.5, 0.5), sc0, st0)
    packed = capi.pack_dna4(dsq)
    sc1, st1 = np.empty(len(w), np.float32), np.empty(len(w), np.int32)
    gpu_ctx.fs_fwd_block_packed4_into(packed, n, w, (0.5, 0.5), sc1, st1)
    assert np.array_equal(st0, st1) and np.array_equal(sc0, sc1)
    gpu_ctx.upload_block_packed4(packed, n)
    sc2, st2 = gpu_ctx.fs_fwd_windows(w)
    assert np.array_equal(st0, st2) and np.array_equal(sc0, sc2)


@pytest.mark.gpu
@pytest.mark.parametrize("entry", ["upload", "block"])
def test_packed_block_feeds_the_orf_finder_and_revcomp(oracle, gpu_ctx, entry):
    """the byte form made on the device from a packed upload is the one a byte upload leaves: ORFs, residues and MSV scores of both
    strands (slot 1 = reverse complement of slot 0) are equal"""
    from bath_b200 import capi, hostapi
    po, lib = oracle, oracle.lib()
    model = hostapi.QueryModel(common.golden("AMP_N.bhmm"))
    rbv, rwv, twv = model.filter_tables()
    gpu_ctx.load_filter_profile(model.filter_params(), rbv, rwv, twv)
    omodel = po.Model(common.golden("AMP_N.bhmm"))
    gpu_ctx.load_fs_profile(3, omodel.rfv(3), omodel.tfv(3))
    rng = np.random.default_rng(5)
    n = 300_001
    dsq = common.random_dna(rng, n, p_degenerate=0.002)
    blocks = np.zeros(1, capi.block_dtype)
    blocks[0]["goff"], blocks[0]["n"], blocks[0]["C"] = 0, n, 0
    gcode = np.frombuffer(C.string_at(lib.bo_gencode_basic(1), 64), np.uint8)
    maxlen = n // 3 + 2
    tjb = np.zeros(maxlen + 1, np.uint8)
    for Lx in range(1, maxlen + 1):
        tjb[Lx] = model.orf_length_params(Lx)[0]
    tjb[0] = tjb[1]
    Ls = np.arange(maxlen + 1)
    p1 = (Ls.astype(np.float32) / (Ls + 1).astype(np.float32)).astype(np.float64)
    null = np.zeros(maxlen + 1, np.float32)
    null[1:] = (Ls[1:] * np.log(p1[1:]) + np.log(1.0 - p1[1:])).astype(np.float32)

    def both_strands():
        out = []
        gpu_ctx.revcomp_slot(0, 1)
        for slot, comp in ((0, 0), (1, 1)):
            gpu_ctx.select_slot(slot)
            per, hits, res = gpu_ctx.orfs_msv_screen(blocks, comp, gcode, 20, tjb, null, -1e30)
            out.append((per.copy(), hits.copy(), res.copy()))
        gpu_ctx.select_slot(0)
        return out

    gpu_ctx.select_slot(0)
    gpu_ctx.upload_block(dsq)
    want = both_strands()
    packed = capi.pack_dna4(dsq)
    gpu_ctx.upload_block(np.full(n + 2, 0, np.uint8))            # overwrite the resident bytes so nothing stale can pass
    if entry == "upload":
        gpu_ctx.upload_block_packed4(packed, n)
    else:
        w = capi.Context.make_windows([1], [300])
        gpu_ctx.fs_fwd_block_packed4_into(packed, n, w, (0.5, 0.5), np.empty(1, np.float32), np.empty(1, np.int32))
    got = both_strands()
    for (p0, h0, r0), (p1_, h1, r1) in zip(want, got):
        assert np.array_equal(p0, p1_) and len(h0) > 1000
        for f in ("block", "index", "start", "end", "n", "frame", "status"):
            assert np.array_equal(h0[f], h1[f]), f
        assert np.array_equal(h0["usc"], h1["usc"])
        for a, b in zip(h0, h1):                                 # residue offsets are handed out in arrival order: compare the residues
            assert np.array_equal(r0[a["offset"]: a["offset"] + a["n"]], r1[b["offset"]: b["offset"] + b["n"]])


@pytest.mark.gpu
def test_segment_upload_equals_one_block(oracle, gpu_ctx):
    """bathgpu_upload_block_segments: pieces that follow one another on the device == the concatenated block uploaded at once
    (Forward scores of windows across the seams bit for bit; reverse complement made from it)"""
    from bath_b200 import capi
    model = oracle.Model(common.golden("AMP_N.bhmm"))
    gpu_ctx.load_fs_profile(3, model.rfv(3), model.tfv(3))
    rng = np.random.default_rng(21)
    lens = [1, 7, 100_003, 8, 250_000, 33_333]
    pieces = [common.random_dna(rng, n, p_degenerate=0.002)[1:-1].copy() for n in lens]
    whole = np.concatenate([[255]] + pieces + [[255]]).astype(np.uint8)
    n = len(whole) - 2
    seams = np.cumsum(lens)[:-1]
    starts = np.concatenate([rng.integers(1, n - 1300, 500), np.maximum(1, seams - 300), [1, n - 1199]])
    w = capi.Context.make_windows(starts, np.full(len(starts), 1200))
    gpu_ctx.select_slot(0)
    gpu_ctx.upload_block(whole)
    sc0, st0 = gpu_ctx.fs_fwd_windows(w)
    gpu_ctx.revcomp_slot(0, 1); gpu_ctx.select_slot(1)
    rc0, _ = gpu_ctx.fs_fwd_windows(w)
    gpu_ctx.select_slot(0)
    gpu_ctx.upload_block(np.zeros(n + 2, np.uint8))
    gpu_ctx.upload_block_segments(pieces)
    sc1, st1 = gpu_ctx.fs_fwd_windows(w)
    gpu_ctx.revcomp_slot(0, 1); gpu_ctx.select_slot(1)
    rc1, _ = gpu_ctx.fs_fwd_windows(w)
    gpu_ctx.select_slot(0)
    assert np.array_equal(sc0, sc1) and np.array_equal(st0, st1) and np.array_equal(rc0, rc1)
    with pytest.raises(capi.BathGpuError):
        gpu_ctx.upload_block_segments([pieces[0], np.zeros(0, np.uint8)])
