"""E-value calibration by simulation: p7_Calibrate with the frameshift branch (src/evalues.c:64-183) and
bathconvert's frameshift-only flow (src/bathconvert.c:128-161).

The golden values are the STATS lines the reference's own programs wrote into the shipped models:
  * AMP_N.bhmm (bathbuild): all five lines -- MSV mu, Viterbi mu, Forward tau, FS3 tau, FS5 tau -- from one generator seeded 42;
  * tRNA-proteins.bhmm (12 models), MET-ct4.bhmm (2 models, codon table 4), PTHR37536.bhmm: the FS3 / FS5 lines of a bathconvert
    run, whose generator is created once (seed 42) and keeps running from one model of the file to the next;
  * tRNA-synthetases.bhmm: the protein lines (written by the program the models were first built with, same algorithm).
Reproducing them to the printed precision pins Easel's fast generator and esl_rnd_FChoose as restated (also used by the
multi-domain branch), the MSV and Viterbi filters, the protein Forward parser and both frameshift Forward recursions against
numbers the reference produced.  Tolerance 2.5e-4: the file gives lambda and the emissions to 5 decimals only.

The CPU tests run the oracle (oracle/calibrate.c) and the product's host library (bath_b200/host/calibrate.cpp) over the CPU
implementation of the stage calls; tests marked gpu run the same host code over libbathgpu.so."""
import pytest

import common

TOL = 2.5e-4


def stats_of(oracle, hmm, index):
    return oracle.Model(common.golden(hmm), index).evparam


def close(got, want):
    return abs(got - want) <= TOL


# ---------------------------------------------------------------- oracle vs the shipped STATS lines

def test_oracle_bathbuild_flow_reproduces_amp_n_stats(oracle):
    m = oracle.Model(common.golden("AMP_N.bhmm"))
    ev, _ = oracle.calibrate(m)                        # lambda as the file gives it
    for z in (0, 2, 4, 6, 7):
        assert close(ev[z], m.evparam[z]), (z, ev[z], m.evparam[z])
    # p7_Lambda from the (rounded) emissions of the file
    assert abs(oracle.lib().bo_Lambda(m.hmm, m.bg) - m.evparam[1]) < 2e-5


@pytest.mark.parametrize("hmm,count", [("tRNA-proteins.bhmm", 12), ("MET-ct4.bhmm", 2), ("PTHR37536.bhmm", 1)])
def test_oracle_bathconvert_flow_reproduces_frameshift_taus(oracle, hmm, count):
    x = 0
    for i in range(count):
        m = oracle.Model(common.golden(hmm), i)
        ev, x = oracle.calibrate(m, which=24, convert_flow=True, rng_x=x)
        assert close(ev[6], m.evparam[6]) and close(ev[7], m.evparam[7]), (hmm, i, ev[6:], m.evparam[6:])


@pytest.mark.parametrize("index,which", [(0, (2, 4)), (1, (0, 2, 4)), (2, (0, 2, 4))])
def test_oracle_protein_stats_of_trna_synthetases(oracle, index, which):
    # model 0's MSV line (-10.1387 in the file, -10.1327 from the file's own numbers) is explained by the test below
    m = oracle.Model(common.golden("tRNA-synthetases.bhmm"), index)
    ev, _ = oracle.calibrate(m, which=7)
    for z in which:
        assert close(ev[z], m.evparam[z]), (index, z, ev[z], m.evparam[z])


def test_trna_synthetases_model0_msv_mu_is_one_byte_cost_on_a_rounding_boundary(oracle):
    """The one shipped STATS value the restatement does not reproduce, bisected: the file stores emissions as -ln p
    with 5 decimals, and the MSV byte cost of residue D at node 137, roundf(-scale_b * log(p / f)), comes out at 0.49999 -- 1.4e-5
    below the rounding boundary, inside the 2.2e-5 the file's rounding leaves open.  The program that wrote the STATS line held the
    unrounded emission: with that one cost at 1 instead of 0 the fit over the same 200 sequences gives the file's value.  Every other
    byte cost, the generator, the sequences and the fit are therefore the reference's (the Viterbi and Forward lines match as is)."""
    import ctypes as C
    import numpy as np
    L = oracle.lib()
    m = oracle.Model(common.golden("tRNA-synthetases.bhmm"), 0)

    def msv_mu():
        out = (C.c_double * 8)()
        x = C.c_uint32(0)
        assert L.bo_Calibrate(m.hmm, m.bg, m.om, m.om_fs3, m.om_fs5, m.ct, 42, float(m.evparam[5]), 1, 0, C.byref(x), out) == 0
        return out[0]

    assert abs(msv_mu() - (-10.1327)) < 6e-5
    h = m.hmm.contents
    p = float(np.ctypeslib.as_array(h.mat, shape=(h.M + 1, 20))[137, 2])
    f_D = float(np.ctypeslib.as_array(m.bg.contents.f, shape=(20,))[2])
    cost = -(3.0 / np.log(2.0)) * np.log(p / f_D)
    assert 0.5 - cost < 2.2e-5 and cost < 0.5                 # on the boundary within the file's rounding of -ln p (5e-6 x scale_b)
    rbv = m.om_tables()[0]
    rbv[2, 137] += 1
    assert abs(msv_mu() - m.evparam[0]) < 6e-5                # -10.1387, the file's value
    rbv[2, 137] -= 1


# ---------------------------------------------------------------- the product's host code over the CPU stage calls

def host_calibrate(be, hmm, index=0, **kw):
    from bath_b200 import hostapi
    model = hostapi.QueryModel(common.golden(hmm), index)
    return hostapi.calibrate(model, backend=be, **kw)


def test_host_library_bathbuild_flow_cpu_backend(oracle):
    be, keep = oracle.cpu_backend(4)
    want = stats_of(oracle, "AMP_N.bhmm", 0)
    ev, _ = host_calibrate(be, "AMP_N.bhmm", lam=want[1])
    for z in (0, 2, 4, 6, 7):
        assert close(ev[z], want[z]), (z, ev[z], want[z])
    ref, _ = oracle.calibrate(oracle.Model(common.golden("AMP_N.bhmm")))
    assert max(abs(a - b) for a, b in zip(ev, ref)) < 1e-6          # same draws, same scores, same fits
    ev2, _ = host_calibrate(be, "AMP_N.bhmm")                        # lambda from p7_Lambda
    assert abs(ev2[1] - want[1]) < 2e-5 and close(ev2[6], want[6])
    del keep


def test_host_library_bathconvert_flow_cpu_backend(oracle):
    be, keep = oracle.cpu_backend(4)
    x = 0
    for i in range(4):
        want = stats_of(oracle, "tRNA-proteins.bhmm", i)
        ev, x = host_calibrate(be, "tRNA-proteins.bhmm", i, convert_flow=True, rng_state=x)
        assert ev[0] == -99999.0 and close(ev[6], want[6]) and close(ev[7], want[7]), (i, ev, want)
    del keep


# ---------------------------------------------------------------- the product path: the same host code over libbathgpu.so

@pytest.mark.gpu
def test_gpu_bathbuild_flow_reproduces_amp_n_stats(oracle, gpu_ctx):
    from bath_b200 import hostapi
    want = stats_of(oracle, "AMP_N.bhmm", 0)
    model = hostapi.QueryModel(common.golden("AMP_N.bhmm"))
    ev, _ = hostapi.calibrate(model, gpu_ctx=gpu_ctx, lam=want[1])
    for z in (0, 2, 4, 6, 7):
        assert close(ev[z], want[z]), (z, ev[z], want[z])
    ref, _ = oracle.calibrate(oracle.Model(common.golden("AMP_N.bhmm")))
    assert ev[0] == pytest.approx(ref[0], abs=1e-9) and ev[2] == pytest.approx(ref[2], abs=1e-9)   # integer filters: bit-exact scores
    assert max(abs(a - b) for a, b in zip(ev, ref)) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("hmm,count", [("tRNA-proteins.bhmm", 12), ("MET-ct4.bhmm", 2), ("PTHR37536.bhmm", 1)])
def test_gpu_bathconvert_flow_reproduces_frameshift_taus(oracle, gpu_ctx, hmm, count):
    from bath_b200 import hostapi
    x = 0
    for i in range(count):
        want = stats_of(oracle, hmm, i)
        model = hostapi.QueryModel(common.golden(hmm), i)
        ev, x = hostapi.calibrate(model, gpu_ctx=gpu_ctx, convert_flow=True, rng_state=x)
        assert close(ev[6], want[6]) and close(ev[7], want[7]), (hmm, i, ev[6:], want[6:])
