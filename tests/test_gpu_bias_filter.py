"""GPU parity: the bias-composition filter (a5) through the C ABI vs the CPU oracle -- BIT-EXACT.

bathgpu_bias_forward (csrc/bias_filter.cuh) against oracle/cpu_backend.c's restatement of esl_hmm_Forward over the filter HMM of
p7_bg_SetFilter (src/p7_bg.c:449-573): ORFs (p7_bg_FilterScore) and the three frames of DNA windows (p7_bg_fs_FilterScore),
several emission tables (model composition + local compositions), ragged lengths, degenerate residues and nucleotides, stop codons.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
KP = 29


def filter_tables(rng, ntab):
    """[ntab][29][2] emission odds as esl_hmm_Configure leaves them: state 0 = background (odds 1), state 1 = a composition"""
    f = rng.dirichlet(np.full(20, 8.0)).astype(np.float32)
    t = np.ones((ntab, KP, 2), np.float32)
    for z in range(ntab):
        compo = rng.dirichlet(np.full(20, 2.0 + 3.0 * z)).astype(np.float32)
        t[z, :20, 0] = f / f
        t[z, :20, 1] = compo / f
        for x, members in ((21, (2, 11)), (22, (7, 9)), (23, (3, 13)), (24, (8,)), (25, (1,)), (26, tuple(range(20)))):
            t[z, x, 1] = compo[list(members)].sum() / f[list(members)].sum()
    return t


def cpu_bias(po, be, kind, items, tables, t10, t11, gcode):
    L = po.lib()
    L.bo_backend_bias_forward.restype = C.c_int
    L.bo_backend_bias_forward.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_float, C.c_float,
                                          C.POINTER(C.c_uint8), C.POINTER(C.c_float)]
    out = np.zeros(len(items) * (3 if kind else 1), np.float32)
    g = np.ascontiguousarray(gcode, np.uint8)
    st = L.bo_backend_bias_forward(be.ctx, kind, items.ctypes.data_as(C.c_void_p), len(items), tables.ctypes.data_as(C.POINTER(C.c_float)),
                                   len(tables), t10, t11, g.ctypes.data_as(C.POINTER(C.c_uint8)), out.ctypes.data_as(C.POINTER(C.c_float)))
    assert st == 0
    return out.reshape(-1, 3) if kind else out


def test_orf_bias_scores_bit_exact(oracle, gpu_ctx):
    from bath_b200 import capi
    rng = np.random.default_rng(11)
    n = 3000
    lens = rng.integers(1, 900, n)
    lens[:5] = (1, 2, 3, 20, 2500)
    res = rng.integers(0, 20, int(lens.sum())).astype(np.uint8)
    res[rng.integers(0, len(res), 400)] = rng.integers(20, 27, 400)          # degenerate / X residues
    off = np.concatenate([[0], np.cumsum(lens)[:-1]])
    tables = filter_tables(rng, 7)
    items = np.zeros(n, capi.bias_item_dtype)
    items["start"], items["L"], items["table"] = off, lens, rng.integers(0, 7, n)
    items["t00"] = (lens.astype(np.float32) / (lens + 1).astype(np.float32)).astype(np.float32)
    t10, t11 = np.float32(1.0) / np.float32(25.0), np.float32(24.0) / np.float32(25.0)

    be, keep = oracle.cpu_backend(4)
    L = oracle.lib()
    L.bo_backend_upload_orfs.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_int64]
    assert L.bo_backend_upload_orfs(be.ctx, res.ctypes.data_as(C.POINTER(C.c_uint8)), len(res)) == 0
    want = cpu_bias(oracle, be, 0, items, tables, t10, t11, np.zeros(64, np.uint8))

    gpu_ctx.upload_orfs(res)
    got = gpu_ctx.bias_forward(0, items, tables, t10, t11)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.abs(got - want).max()
    assert np.isfinite(got).all() and (got != 0).any()


def test_window_frame_bias_scores_bit_exact(oracle, gpu_ctx):
    from bath_b200 import capi
    rng = np.random.default_rng(12)
    N = 60000
    dsq = np.full(N + 2, 255, np.uint8)
    dsq[1:-1] = rng.integers(0, 4, N)
    dsq[1 + rng.integers(0, N, 300)] = rng.integers(4, 16, 300)              # degenerate nucleotides: their codons are skipped
    gcode = rng.integers(0, 20, 64).astype(np.uint8)                         # any codon -> residue map will do
    gcode[[48, 50, 56]] = 27                                                 # three stop codons
    n = 800
    L = rng.integers(0, 3000, n)
    L[:6] = (0, 1, 2, 3, 4, 5)
    start = np.array([rng.integers(1, N - l + 2) for l in L])
    start[6], L[6] = 1, N                                                    # the whole sequence
    tables = filter_tables(rng, 5)
    items = np.zeros(n, capi.bias_item_dtype)
    items["start"], items["L"], items["table"] = start, L, rng.integers(0, 5, n)
    la = (L // 3)
    items["t00"] = (la.astype(np.float32) / (la + 1).astype(np.float32)).astype(np.float32)
    t10, t11 = np.float32(1.0) / np.float32(17.0), np.float32(16.0) / np.float32(17.0)

    be, keep = oracle.cpu_backend(4)
    lib = oracle.lib()
    lib.bo_backend_upload_block.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_int64]
    assert lib.bo_backend_upload_block(be.ctx, dsq.ctypes.data_as(C.POINTER(C.c_uint8)), N) == 0
    want = cpu_bias(oracle, be, 1, items, tables, t10, t11, gcode)

    gpu_ctx.select_slot(0)
    gpu_ctx.upload_block(dsq)
    got = gpu_ctx.bias_forward(1, items, tables, t10, t11, gcode)
    assert got.shape == (n, 3)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.abs(got - want).max()
    assert (got[:3] == 0).all()                                              # no whole codon: the recursion is not entered


def test_bias_forward_rejects_items_outside_the_resident_data(gpu_ctx):
    from bath_b200 import capi
    items = np.zeros(1, capi.bias_item_dtype)
    items["start"], items["L"] = 10 ** 12, 50
    with pytest.raises(capi.BathGpuError):
        gpu_ctx.bias_forward(0, items, np.ones((1, KP, 2), np.float32), 0.1, 0.9)
