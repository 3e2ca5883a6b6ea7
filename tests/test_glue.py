"""glue/ehmm_glue.cpp -- the bodies src/*.cpp would get -- compiled for real.  R and Rcpp are not in this image,
so the glue is built against tests/mock_rcpp/Rcpp.h (the slice of the Rcpp API it uses, on plain containers).
CPU: it compiles, links against libehmm_b200.so and defines the reference's thirteen exports.  GPU:
tests/glue_harness.cpp drives every export and the results equal the ctypes path of the same library."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO_DIR = os.path.join(ROOT, "epigrahmm_b200")
CXX = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "mock_rcpp"), "-I", os.path.join(ROOT, "include")]

# the .Call table of the reference (src/RcppExports.cpp:174-194) minus simulateMarkovChain (a documentation utility)
REFERENCE_EXPORTS = ["aggregate", "computeBIC", "computeQFunction", "computeViterbiSequence", "consensusRejectionControlled",
                     "differentialRejectionControlled", "expStep", "getMarginalProbability", "innerMaxStepProb", "maxStepProb",
                     "rbinomVectorized", "reweight"]


def _build(tmp_path, ehmm):
    glue_o, har_o, exe = (str(tmp_path / n) for n in ("glue.o", "harness.o", "glue_harness"))
    subprocess.run(CXX + ["-c", os.path.join(ROOT, "glue", "ehmm_glue.cpp"), "-o", glue_o], check=True)
    subprocess.run(CXX + ["-c", os.path.join(ROOT, "tests", "glue_harness.cpp"), "-o", har_o], check=True)
    subprocess.run(["g++", "-o", exe, har_o, glue_o, "-L", SO_DIR, "-lehmm_b200", "-Wl,-rpath," + SO_DIR], check=True)
    return glue_o, exe


def test_glue_compiles_links_and_defines_the_reference_exports(tmp_path, ehmm):
    glue_o, exe = _build(tmp_path, ehmm)
    syms = subprocess.run(["nm", "-C", "--defined-only", glue_o], capture_output=True, text=True, check=True).stdout
    for name in REFERENCE_EXPORTS:
        assert re.search(r" T %s\(" % name, syms), name
    # the reference's registrations carry these names and arities (src/RcppExports.cpp:174-194)
    ref = os.path.join("/root/reference", "src", "RcppExports.cpp")
    if os.path.exists(ref):
        table = open(ref).read()
        for name in REFERENCE_EXPORTS:
            assert "_epigraHMM_%s" % name in table, name
    und = subprocess.run(["nm", "-u", glue_o], capture_output=True, text=True, check=True).stdout
    used = set(re.findall(r"\b(ehmm_[a-z0-9_]+)", und))
    header = open(os.path.join(ROOT, "include", "ehmm.h")).read()
    assert used and all(re.search(r"\b%s\(" % u, header) for u in used), used


@pytest.mark.gpu
@pytest.mark.skipif(not has_gpu(), reason="needs a CUDA device")
def test_glue_exports_equal_the_ctypes_path(tmp_path, ehmm, oracle):
    from epigrahmm_b200 import synth
    from epigrahmm_b200.rng import RState
    _, exe = _build(tmp_path, ehmm)
    M, N = 20_011, 3
    d = synth.make_consensus(M, N, seed=8)
    th = synth.THETA_CONSENSUS
    seed = np.asarray(RState(7).state, dtype=np.uint32)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.txt")
    with open(fin, "wb") as f:
        f.write(struct.pack("qq", M, N))
        f.write(np.asfortranarray(d["counts"], dtype=np.int32).tobytes(order="F"))
        f.write(np.asfortranarray(d["offsets"], dtype=np.float64).tobytes(order="F"))
        f.write(seed.tobytes())
        f.write(np.asarray(th["pi"], dtype=np.float64).tobytes())
        f.write(np.asfortranarray(th["gamma"], dtype=np.float64).tobytes(order="F"))
        f.write(np.asarray(th["psi"][0], dtype=np.float64).tobytes())
        f.write(np.asarray(th["psi"][1], dtype=np.float64).tobytes())
        # the differential block: 3 conditions x 1 replicate, 6 mixture patterns
        M2 = 15_013
        dd = synth.make_differential(M2, 3, 1, seed=9)
        td = synth.THETA_DIFFERENTIAL
        Bd = dd["B"]
        delta = np.full(Bd, 1.0 / Bd)
        f.write(struct.pack("qqq", M2, 3, Bd))
        f.write(np.asfortranarray(dd["counts"], dtype=np.int32).tobytes(order="F"))
        f.write(np.asfortranarray(dd["offsets"], dtype=np.float64).tobytes(order="F"))
        f.write(np.asfortranarray(dd["design"].astype(np.int32)).tobytes(order="F"))
        f.write(delta.tobytes())
        f.write(np.asarray(td["psi"], dtype=np.float64).tobytes())
        f.write(np.asarray(td["pi"], dtype=np.float64).tobytes())
        f.write(np.asfortranarray(td["gamma"], dtype=np.float64).tobytes(order="F"))
    r = subprocess.run([exe, fin, fout], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = {}
    for ln in open(fout):
        t = ln.split()
        got[t[0]] = np.array(t[2:], dtype=np.float64)
        assert got[t[0]].size == int(t[1])
    assert "FAILED" not in got and got["error_raised"][0] == 1

    with ehmm.Session("glue_ref", M, 2) as s:
        s.set_data(d["counts"], d["offsets"])
        G = s.build_groups()
        s.rng_set_state(seed)
        s.rc_hint(0.05)
        s.expstep_mode(1)
        s.loglik_consensus(th["psi"], 2.2250738585072014e-308)
        s.expstep(th["pi"], th["gamma"])
        m = s.maxstepprob()
        s.rc_consensus(0.05)
        assert got["G"][0] == G
        assert np.array_equal(got["pi"], m["pi"]) and np.array_equal(got["gamma"], np.ravel(m["gamma"], order="F"))
        for k in range(2):
            b = s.rc_buckets(k)
            ids = np.flatnonzero(b)
            tab = got["rc%d" % k].reshape(-1, 2, order="F")      # (sum, id) rows, ascending id: src/aggregate.cpp:32-44
            assert np.array_equal(tab[:, 1], ids) and np.array_equal(tab[:, 0], b[ids])
        after = np.asarray(s.rng_get_state(), dtype=np.uint32)
        assert got["seed_after"][0] == 10403 and np.array_equal(got["seed_after"][1:].astype(np.uint32), after)
        v, gr = s.glm_nb(1, [th["psi"][1][0], 1.0 / th["psi"][1][1]])
        assert got["glm_value"][0] == v and np.array_equal(got["glm_grad"], gr)
        assert got["q"][0] == s.qfunction(th["pi"], th["gamma"]) and got["bic"][0] == s.bic(5, N)
        assert np.array_equal(got["viterbi"], s.viterbi(th["pi"], th["gamma"]))
        assert np.array_equal(got["marginal"], np.ravel(s.marginal_probability(), order="F"))
        lfp = np.ravel(s.fetch("logFP"), order="F")
        assert np.array_equal(got["logFP"], lfp) and np.array_equal(got["logFP_b"], lfp)
        assert np.array_equal(got["peaks"].astype(bool), s.call_peaks(0.05))

    with ehmm.Session("glue_ref_d", M2, 3) as s:
        s.set_data(dd["counts"], dd["offsets"])
        s.set_design(dd["design"])
        Gd = s.build_groups()
        s.rng_set_state(seed)
        s.rc_hint(0.05)
        s.expstep_mode(1)
        s.loglik_differential_hmm(delta, td["psi"], 2.2250738585072014e-308)
        s.expstep(td["pi"], td["gamma"])
        md = s.maxstepprob()
        s.inner_expstep(delta, td["psi"], 2.2250738585072014e-308)
        dl = s.inner_maxstepprob()
        s.rc_differential(0.05)
        assert got["d_G"][0] == Gd
        assert np.array_equal(got["d_gamma"], np.ravel(md["gamma"], order="F")) and np.array_equal(got["d_delta"], dl)
        for c in range(Bd + 2):
            b = s.rc_buckets(c)
            ids = np.flatnonzero(b)
            tab = got["d_rc%d" % c].reshape(-1, 2, order="F")
            assert np.array_equal(tab[:, 1], ids) and np.array_equal(tab[:, 0], b[ids]), c
        psi = td["psi"]
        v, gr = s.glm_mix_nb([psi[0], psi[2], psi[1], psi[3]], 2.2250738585072014e-308)
        assert got["d_glm_value"][0] == v and np.array_equal(got["d_glm_grad"], gr)
        assert np.array_equal(got["d_mixtureProb"], np.ravel(s.fetch("mixtureProb"), order="F"))

    # the stand-alone helpers, with the mock generator's uniforms replayed here
    x = ((np.arange(M) * 37) % 101) / 100.0
    grp = 1 + (np.arange(M) * 7) % 13
    agg = oracle.aggregate(x, grp.astype(np.float64))
    assert np.array_equal(got["aggregate"].reshape(-1, 2, order="F"), agg)

    def draws(n, state=42):
        out = np.empty(n)
        mask = (1 << 64) - 1
        for i in range(n):
            state ^= (state << 13) & mask
            state ^= state >> 7
            state ^= (state << 17) & mask
            out[i] = ((state >> 11) + 0.5) / 9007199254740992.0
        return out
    p = 0.3
    need = int(np.sum((x < p) & (x / p != 0.0)))
    want, used = oracle.reweight(x, p, uniforms=draws(need))
    assert used == need and np.array_equal(got["reweight"], want)
    u = iter(draws(M))
    rb = np.empty(M)
    for j, pp in enumerate(x):                      # R::rbinom(1, pp): no draw for 0 or 1
        if pp == 0.0 or pp == 1.0:
            rb[j] = pp
        else:
            q = 1.0 - min(pp, 1.0 - pp)
            ix = 0 if next(u) < q else 1
            rb[j] = 1 - ix if pp > 0.5 else ix
    assert np.array_equal(got["rbinom"], rb)
