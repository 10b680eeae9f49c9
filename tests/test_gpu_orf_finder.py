"""GPU parity: six-frame translation, MSV and the F1 screen in one device call (bathgpu_orfs_msv_screen) against
the oracle's ORF finder (oracle/orfs.c) and MSV filter -- EXACT: the same ORFs in the same order with the same ranks, the same
residues, bit-identical MSV scores, the same survivors.  Blocks with overlap context on both strands, degenerate nucleotides,
blocks too short to search, two genetic codes."""
import ctypes as C

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def setup(po, gpu_ctx, hmmfile, index=0):
    from test_gpu_orf_filters import host_params
    model = po.Model(common.golden(hmmfile), index)
    rbv, rwv, twv, _, _ = model.om_tables()
    gpu_ctx.load_filter_profile(host_params(po, model, 16, 8), rbv, rwv, twv)
    return model


def tables(po, model, maxlen):
    lib = po.lib()
    o = model.om.contents
    tjb, null = np.zeros(maxlen + 1, np.uint8), np.zeros(maxlen + 1, np.float32)
    for L in range(1, maxlen + 1):
        lib.bo_oprofile_ReconfigLength(model.om, L)
        tjb[L] = o.tjb_b
        lib.bo_bg_SetLength(model.bg, L)
        null[L] = lib.bo_bg_NullOne(model.bg, L)
    tjb[0] = tjb[1]
    return tjb, null


@pytest.mark.parametrize("ct", [1, 4])
@pytest.mark.parametrize("complement", [0, 1])
def test_orfs_and_msv_match_oracle(oracle, gpu_ctx, ct, complement):
    po, lib = oracle, oracle.lib()
    from bath_b200 import capi
    model = setup(po, gpu_ctx, "AMP_N.bhmm")
    rng = np.random.default_rng(100 * ct + complement)
    n = 50000
    dsq = common.random_dna(rng, n, p_degenerate=0.002)
    mat = common.hmm_mat(model)
    for at in (3000, 21000, 40000):                                   # a few homologs so that some ORFs score high (and one overflows)
        ins = common.sample_homolog(rng, mat, fs_rate=0.0, stop_rate=0.0, sub_from_model=(at != 21000))
        dsq[at: at + len(ins)] = ins
    gpu_ctx.upload_block(dsq)
    # blocks of 12000 with 500 nt of context, plus one block too short to search
    blocks = np.zeros(6, capi.block_dtype)
    pos, b = 0, 0
    while pos < n and b < 5:
        ctx_len = 0 if b == 0 else 500
        blocks[b]["goff"], blocks[b]["n"], blocks[b]["C"] = pos - ctx_len, min(12000 + ctx_len, n - pos + ctx_len), ctx_len
        pos += 12000; b += 1
    blocks[5]["goff"], blocks[5]["n"], blocks[5]["C"] = n - 10, 0, 0  # n < 15: passed as empty
    gcode = np.frombuffer(C.string_at(lib.bo_gencode_basic(ct), 64), np.uint8)
    tjb, null = tables(po, model, 12500 // 3 + 2)
    per, hits, res = gpu_ctx.orfs_msv_screen(blocks, complement, gcode, 20, tjb, null, -1e30)     # min_bits = -inf: every scored ORF survives
    o = model.om.contents
    z = 0
    n_overflow = n_context = 0
    for bi in range(6):
        want = po.find_orfs(dsq[int(blocks[bi]["goff"]):], int(blocks[bi]["n"]), gcode, 20) if blocks[bi]["n"] >= 3 else []
        assert per[bi] == len(want), (bi, per[bi], len(want))
        for idx, (start, end, frame, aa) in enumerate(want):
            Cn, bn = int(blocks[bi]["C"]), int(blocks[bi]["n"])
            in_context = ((bn - start + 1) < Cn) if complement else (end < Cn)
            if in_context:
                n_context += 1
                continue
            h = hits[z]; z += 1
            assert (h["block"], h["index"], h["start"], h["end"], h["frame"], h["n"]) == (bi, idx, start, end, frame, len(aa)), (bi, idx, h)
            assert np.array_equal(res[h["offset"]: h["offset"] + h["n"]], aa)
            d = np.concatenate([[255], aa, [255]]).astype(np.uint8)
            lib.bo_oprofile_ReconfigLength(model.om, len(aa))
            sc = C.c_float()
            st = lib.bo_MSVFilter(po.u8ptr(d), len(aa), model.om, C.byref(sc))
            assert h["status"] == st and (h["usc"] == sc.value or (np.isinf(h["usc"]) and np.isinf(sc.value))), (bi, idx, h, sc.value)
            n_overflow += st == 16
    assert z == len(hits) and z > 500 and n_context > 0
    print(f"ct={ct} complement={complement}: {int(per.sum())} ORFs in 6 blocks, {n_context} inside context, {n_overflow} MSV overflows")


def test_screen_keeps_what_can_pass(oracle, gpu_ctx):
    """with a real threshold the survivors are exactly the scored ORFs at or above it (or overflowed), in order"""
    po, lib = oracle, oracle.lib()
    from bath_b200 import capi
    model = setup(po, gpu_ctx, "PTH2.bhmm")
    rng = np.random.default_rng(5)
    n = 200000
    dsq = common.random_dna(rng, n)
    gpu_ctx.upload_block(dsq)
    blocks = np.zeros(1, capi.block_dtype)
    blocks[0]["goff"], blocks[0]["n"], blocks[0]["C"] = 0, n, 0
    gcode = np.frombuffer(C.string_at(lib.bo_gencode_basic(1), 64), np.uint8)
    tjb, null = tables(po, model, 2000)
    per_all, all_hits, _ = gpu_ctx.orfs_msv_screen(blocks, 0, gcode, 20, tjb, null, -1e30)
    bits_all = (all_hits["usc"].astype(np.float64) - null[np.minimum(all_hits["n"], 2000)].astype(np.float64)) / 0.69314718055994529
    cut = float(np.quantile(bits_all, 0.95))
    per, hits, res = gpu_ctx.orfs_msv_screen(blocks, 0, gcode, 20, tjb, null, cut)
    assert per[0] == per_all[0] == len(all_hits)
    bits = (all_hits["usc"].astype(np.float64) - null[np.minimum(all_hits["n"], 2000)].astype(np.float64)) / 0.69314718055994529
    keep = (all_hits["status"] != 0) | (bits >= cut)
    assert 0 < keep.sum() < len(all_hits) // 5
    assert np.array_equal(hits["index"], all_hits["index"][keep]) and np.array_equal(hits["usc"], all_hits["usc"][keep])
    assert int(hits["n"].sum()) == len(res)


def test_orf_finder_argument_errors(oracle, gpu_ctx):
    from bath_b200 import capi
    model = setup(oracle, gpu_ctx, "AMP_N.bhmm")
    dsq = common.random_dna(np.random.default_rng(0), 1000)
    gpu_ctx.upload_block(dsq)
    blocks = np.zeros(1, capi.block_dtype)
    blocks[0]["goff"], blocks[0]["n"] = 500, 600                       # runs past the uploaded sequence
    with pytest.raises(capi.BathGpuError) as e:
        gpu_ctx.orfs_msv_screen(blocks, 0, np.zeros(64, np.uint8), 20, np.zeros(10, np.uint8), np.zeros(10, np.float32), 0.0)
    assert e.value.code == capi.EINVAL


def test_orfs_spanning_many_tiles(oracle, gpu_ctx):
    """stop-free stretches longer than the ORF pass's tile of 2048 positions (a poly-codon run, a run of N, and random DNA between them):
    the previous stop of a frame then lies several tiles back, the path on which the kernel still walks over the class bytes"""
    po, lib = oracle, oracle.lib()
    from bath_b200 import capi
    model = setup(po, gpu_ctx, "AMP_N.bhmm")
    rng = np.random.default_rng(77)
    n = 60000
    dsq = common.random_dna(rng, n)
    dsq[5001:5001 + 9000] = np.tile(np.array([2, 1, 0], np.uint8), 3000)        # GCA x 3000: no stop in any frame for 9 kb
    dsq[20001:20001 + 7000] = 15                                                # N x 7000: every codon translates to X, no stop
    dsq[40001:40001 + 4100] = np.tile(np.array([0, 0, 2, 1], np.uint8), 1025)   # period 4: AAG CAA GCA AGC ..., stop-free
    gpu_ctx.upload_block(dsq)
    blocks = np.zeros(2, capi.block_dtype)
    blocks[0]["goff"], blocks[0]["n"], blocks[0]["C"] = 0, 33000, 0
    blocks[1]["goff"], blocks[1]["n"], blocks[1]["C"] = 33000 - 600, n - 33000 + 600, 600
    gcode = np.frombuffer(C.string_at(lib.bo_gencode_basic(1), 64), np.uint8)
    tjb, null = tables(po, model, 33000 // 3 + 2)
    for complement in (0, 1):
        per, hits, res = gpu_ctx.orfs_msv_screen(blocks, complement, gcode, 20, tjb, null, -1e30)
        z = 0
        longest = 0
        for bi in range(2):
            want = po.find_orfs(dsq[int(blocks[bi]["goff"]):], int(blocks[bi]["n"]), gcode, 20)
            assert per[bi] == len(want), (complement, bi, per[bi], len(want))
            for idx, (start, end, frame, aa) in enumerate(want):
                Cn, bn = int(blocks[bi]["C"]), int(blocks[bi]["n"])
                if ((bn - start + 1) < Cn) if complement else (end < Cn):
                    continue
                h = hits[z]; z += 1
                assert (h["block"], h["index"], h["start"], h["end"], h["frame"], h["n"]) == (bi, idx, start, end, frame, len(aa)), (bi, idx, h)
                assert np.array_equal(res[h["offset"]: h["offset"] + h["n"]], aa)
                longest = max(longest, len(aa))
        assert z == len(hits)
        assert longest > 2800                        # the 9 kb run is one ORF per frame
