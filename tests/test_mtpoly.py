"""The GF(2) jump-ahead arithmetic behind the parallel Mersenne-Twister (csrc/mtpoly.h) is plain
host C++: compile it with g++ and check jumped windows against the sequentially advanced stream."""
import os
import subprocess
import tempfile

from conftest import ROOT

SRC = r'''
#include "mtpoly.h"
#include <stdio.h>
using namespace mtpoly;
int main() {
    Poly phi = min_poly();
    if (phi.empty()) { printf("BAD degree\n"); return 1; }
    int wt = 0; for (int i = 0; i <= DEG; i++) wt += get(phi, i);
    printf("weight %d\n", wt);
    const int log2J = 12;
    auto polys = jump_polys(log2J, 4);
    RawMT g; uint32_t s = 4357; for (int i = 0; i < MT_N; i++) { s = 69069u * s + 1u; g.mt[i] = s; }
    const long J = 1L << log2J, total = 1 + (J << 3) + DEG + 2 * MT_N;
    std::vector<uint32_t> w(total);
    for (int i = 0; i < MT_N; i++) w[i] = g.mt[i];
    for (long i = MT_N; i < total; i++) w[i] = g.next();
    int bad = 0;
    for (int k = 0; k < 4; k++) {
        uint32_t out[MT_N];
        apply(polys[k], &w[1], out);
        for (int p = 0; p < MT_N; p++) bad += out[p] != w[1 + (J << k) + p];
    }
    printf("mismatches %d\n", bad);
    return bad != 0;
}
'''


def test_jump_polynomials_match_sequential_stream():
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.cpp")
        open(src, "w").write(SRC)
        exe = os.path.join(td, "t")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "epigrahmm_b200", "csrc"), "-o", exe, src])
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "weight 135" in out.stdout      # MT19937's characteristic polynomial has 135 terms
        assert "mismatches 0" in out.stdout
