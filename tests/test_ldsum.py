"""R's cumsum (long double accumulation, doubles stored) without a long double: csrc/ldsum.h compiled for the
host against the compiler's own 80-bit type, and the oracle's fdrControl against a numpy longdouble twin."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(np.finfo(np.longdouble).nmant != 63, reason="needs the x87 extended type")
def test_emulated_long_double_sum_is_bit_exact(tmp_path):
    exe = str(tmp_path / "ldsum_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "native", "ldsum_check.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    bad, n = map(int, r.stdout.split())
    assert r.returncode == 0 and bad == 0 and n == 2_000_000


def _twin(prob, fdr):
    notpp = 1.0 - prob
    order = np.argsort(notpp, kind="stable")
    cs = np.cumsum(notpp[order].astype(np.longdouble)).astype(np.float64)
    out = np.zeros(prob.size, dtype=bool)
    out[order] = cs / np.arange(1, prob.size + 1) < fdr
    return out


@pytest.mark.skipif(np.finfo(np.longdouble).nmant != 63, reason="needs the x87 extended type")
def test_oracle_fdr_control(oracle):
    rng = np.random.default_rng(3)
    # the reference's own known-answer test (tests/testthat/test-base.R:1-4): sum(fdrControl(seq(0, 1, 0.001), 0.05)) == 101
    assert oracle.fdrControl(np.arange(1001) * 0.001, 0.05).sum() == 101
    assert list(oracle.fdrControl(np.array([0.99, 0.5, 0.01]), 0.05)) == [True, False, False]
    for n, fdr in ((1, 0.05), (17, 0.2), (5000, 0.05), (200_000, 0.01)):
        prob = np.where(rng.random(n) < 0.1, 1.0 - rng.random(n) * 1e-3, rng.random(n) ** 4)
        prob[rng.integers(0, n, n // 10)] = 1.0          # ties at notpp = 0
        prob[rng.integers(0, n, n // 10)] = 0.75         # ties inside the sort
        got = oracle.fdrControl(prob, fdr)
        assert np.array_equal(got, _twin(prob, fdr))
    with pytest.raises(ValueError):
        oracle.fdrControl(np.array([0.5, 1.5]), 0.05)
    sys.path.insert(0, ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("_em_host", os.path.join(ROOT, "epigrahmm_b200", "em.py"))
    em = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(em)
    prob = rng.random(3000) ** 3
    assert np.array_equal(em.fdrControl(prob, 0.1), oracle.fdrControl(prob, 0.1))
