"""CPU-side checks (no GPU): the C-ABI libraries load and export every symbol their headers declare, the
host-side model set-up is bit-identical to the oracle's, and the synthetic-workload generators are seeded."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import common

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"\w+)\s*\(", text)))


def test_libbathgpu_exports_every_declared_symbol():
    from bath_b200 import build, capi
    build.build_library()
    lib = C.CDLL(build.library_path())
    names = declared_functions("bathgpu.h", "bathgpu_")
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bathgpu.h but not exported"
    assert sorted(capi.EXPORTS) == names, "capi.EXPORTS must list exactly the header's functions"


def test_libbathhost_exports_every_declared_symbol():
    from bath_b200 import build, hostapi
    build.build_host_library()
    lib = C.CDLL(build.host_library_path())
    names = declared_functions("bathhost.h", "bathhost_")
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bathhost.h but not exported"
    assert sorted(hostapi.EXPORTS) == names


def test_no_device_is_an_error_not_a_fallback():
    """Without a CUDA device the product refuses to run (BATHGPU_ENODEVICE); it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bath_b200 import capi
    with pytest.raises(capi.BathGpuError) as e:
        capi.Context(0)
    assert e.value.code == capi.ENODEVICE


@pytest.mark.parametrize("hmmfile,index", [("AMP_N.bhmm", 0), ("2OG-FeII_Oxy_3.bhmm", 0), ("PTH2.bhmm", 0),
                                           ("tRNA-synthetases.bhmm", 0), ("tRNA-synthetases.bhmm", 1),
                                           ("tRNA-synthetases.bhmm", 2), ("PTHR37536.bhmm", 0)])
def test_host_model_setup_is_bit_identical_to_oracle(oracle, hmmfile, index):
    from bath_b200 import hostapi
    o = oracle.Model(common.golden(hmmfile), index)
    h = hostapi.QueryModel(common.golden(hmmfile), index)
    assert h.M == o.M
    if o.max_length > 0:
        assert h.max_length == o.max_length
    # p7_Builder_MaxLength restated: reproduces the MAXL line of every shipped model that has one (known-answer test),
    # and supplies it for the models that lack one, as bathsearch does (src/bathsearch.c:761-762)
    assert h.lib.bathhost_model_computed_max_length(h.h) == h.max_length
    assert h.evparam == o.evparam
    for which in (3, 5):
        assert np.array_equal(h.rfv(which).view(np.uint32), o.rfv(which).view(np.uint32))
        assert np.array_equal(h.tfv(which).view(np.uint32), o.tfv(which).view(np.uint32))
        gm = (o.gm_fs3 if which == 3 else o.gm_fs5).contents
        mc = gm.maxcodons
        oc = np.ctypeslib.as_array(gm.codons, shape=((o.M + 1) * (mc + 1),))[: (o.M + 1) * mc].reshape(o.M + 1, mc)
        oi = np.ctypeslib.as_array(gm.indel_pos, shape=((o.M + 1) * (mc + 1),))[: (o.M + 1) * mc].reshape(o.M + 1, mc)
        assert np.array_equal(h.codons(which), oc)
        assert np.array_equal(h.indel_pos(which), oi)
    assert np.array_equal(h.mat(), common.hmm_mat(o))


def test_model_file_errors():
    from bath_b200 import hostapi
    with pytest.raises(IOError):
        hostapi.QueryModel(common.golden("no-such-file.bhmm"))
    with pytest.raises(IOError):
        hostapi.QueryModel(common.golden("AMP_N.bhmm"), 1)       # single-model file: index 1 is EOF
    with pytest.raises(IOError):
        hostapi.QueryModel(common.golden("target-AMP_N.fa"))     # not a profile file
    assert hostapi.load().bathhost_model_count(os.fsencode(common.golden("tRNA-synthetases.bhmm"))) == 3


def test_length_model_matches_reference_formula():
    """pmove = (2 + nj) / (L + 2 + nj) in float (src/impl_sse/p7_fs_oprofile.c:636-651)"""
    from bath_b200 import capi, hostapi
    for L in (1, 20, 133, 400, 100000):
        for nj in (0.0, 1.0):
            pm, pl = hostapi.length_model(L, nj)
            want = np.float32(np.float32(2.0) + np.float32(nj)) / (np.float32(L) + np.float32(2.0) + np.float32(nj))
            assert pm == want and pl == np.float32(1.0) - want
    w = capi.Context.make_windows([1, 5], [399, 1203], nj=1.0)
    assert w["pmove"][0] == hostapi.length_model(133, 1.0)[0]
    assert w["ploop"][1] == hostapi.length_model(401, 1.0)[1]


def test_synthetic_genome_is_seeded_and_planted():
    from bath_b200 import hostapi, synth
    m = hostapi.QueryModel(common.golden("AMP_N.bhmm"))
    a, pa = synth.planted_genome(np.random.default_rng(42), 200000, m.mat(), every=50000)
    b, pb = synth.planted_genome(np.random.default_rng(42), 200000, m.mat(), every=50000)
    c, _ = synth.planted_genome(np.random.default_rng(43), 200000, m.mat(), every=50000)
    assert np.array_equal(a, b) and pa == pb and not np.array_equal(a, c)
    assert a[0] == 255 and a[-1] == 255 and a[1:-1].max() <= 3 and len(pa) == 4
    s, l = synth.tile_windows(200000, 1200)
    assert s[0] == 1 and s[-1] + l[-1] - 1 == 200000 and np.all(l == 1200)
    s, l = synth.tile_windows(500, 1200)
    assert list(s) == [1] and list(l) == [500]


def test_host_integer_filter_tables_match_oracle(oracle):
    """a1: byte/word score systems of the protein profile (mf_conversion / vf_conversion) and the per-length integers"""
    from bath_b200 import hostapi
    lib = oracle.lib()
    for hmmfile, index in [("AMP_N.bhmm", 0), ("tRNA-synthetases.bhmm", 1), ("PTHR37536.bhmm", 0)]:
        o = oracle.Model(common.golden(hmmfile), index)
        h = hostapi.QueryModel(common.golden(hmmfile), index)
        rbv, rwv, twv, _, _ = o.om_tables()
        hb, hw, ht = h.filter_tables()
        assert np.array_equal(rbv[:, 1:], hb[:, 1:]) and np.array_equal(rwv[:, 1:], hw[:, 1:]) and np.array_equal(twv, ht)
        oc, p = o.om.contents, h.filter_params()
        assert (p["tbm_b"], p["tec_b"], p["base_b"], p["bias_b"], p["base_w"], p["ddbound_w"], p["xw_E_move"], p["xw_E_loop"]) == \
               (oc.tbm_b, oc.tec_b, oc.base_b, oc.bias_b, oc.base_w, oc.ddbound_w, oc.xw[0][0], oc.xw[0][1])
        assert p["scale_b"] == oc.scale_b and p["scale_w"] == oc.scale_w
        for L in (20, 57, 133, 400, 1234, 100000):
            lib.bo_oprofile_ReconfigLength(o.om, L)
            assert h.orf_length_params(L) == (oc.tjb_b, oc.xw[1][0])


def test_msv_shortcut_equivalence(oracle):
    """The GPU runs the J-state MSV recursion only; the reference tries the J-less SSV shortcut first.  Whenever some cell
    beats the begin score the two give the same score and status (src/impl_sse/ssvfilter.c:14-210): checked on the oracle's two
    paths.  (When no cell does -- e.g. an all-X ORF -- the shortcut reports the begin score; the kernel applies that floor, see
    tests/test_gpu_edge_cases.py.)"""
    import ctypes as C
    lib = oracle.lib()
    model = oracle.Model(common.golden("AMP_N.bhmm"))
    mat = common.hmm_mat(model)
    rng = np.random.default_rng(0)
    p = mat[1:] / mat[1:].sum(axis=1, keepdims=True)
    for t in range(120):
        if t % 3 == 0:
            s = np.array([rng.choice(20, p=p[k]) for k in range(model.M)])
            s = np.concatenate([rng.integers(0, 20, int(rng.integers(0, 30))), s[int(rng.integers(0, 60)):]])
        else:
            s = rng.integers(0, 20, int(rng.integers(20, 400)))
        d = np.concatenate([[255], s, [255]]).astype(np.uint8)
        lib.bo_oprofile_ReconfigLength(model.om, len(s))
        a, b = C.c_float(), C.c_float()
        s1 = lib.bo_MSVFilter_opt(oracle.u8ptr(d), len(s), model.om, 1, C.byref(a))
        s2 = lib.bo_MSVFilter_opt(oracle.u8ptr(d), len(s), model.om, 0, C.byref(b))
        assert s1 == s2 and (a.value == b.value or (np.isinf(a.value) and np.isinf(b.value)))
