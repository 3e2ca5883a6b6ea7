"""The C-ABI library loads without a GPU, exports exactly what include/ehmm.h declares, and fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT, has_gpu


def _declared():
    src = open(os.path.join(ROOT, "include", "ehmm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ehmm_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(ehmm):
    out = subprocess.run(["nm", "-D", "--defined-only", ehmm.SO_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (ehmm_[a-z0-9_]+)", out))
    declared = _declared()
    assert len(declared) >= 40
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    extra = sorted(exported - set(declared))
    assert not extra, extra


def test_binding_covers_header(ehmm):
    from epigrahmm_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == _declared()
    assert _lib.lib.ehmm_version() >= 100


def test_library_is_sm100a_only(ehmm):
    out = subprocess.run(["cuobjdump", "--list-elf", ehmm.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(ehmm):
    with pytest.raises(ehmm.EhmmError) as e:
        ehmm.Session("nogpu", 100, 2)
    assert e.value.code == 2  # EHMM_ERR_CUDA
    from epigrahmm_b200 import api
    import numpy as np
    with pytest.raises(ehmm.EhmmError):
        api.expStep([0.5, 0.5], np.full((2, 2), 0.5), np.zeros((10, 2)), "/tmp/x.h5")
    with pytest.raises(ehmm.EhmmError):
        api.maxStepProb("/tmp/never-written.h5")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "epigrahmm_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "libehmm_oracle" not in txt, f


def test_header_is_plain_c_and_links_without_python(ehmm, tmp_path):
    """include/ehmm.h compiles as C99 (no torch / C++ types in the signatures) and a plain C program that
    takes the address of every declared entry point links against the shared library and runs"""
    names = _declared()
    lines = ['#include <stdio.h>', '#include "ehmm.h"', 'typedef void (*anyfn)(void);', 'int main(void) {', '    anyfn fn[] = {']
    lines += ['        (anyfn)%s,' % n for n in names]
    lines += ['    };', '    int n = 0;', '    unsigned i;',
              '    for (i = 0; i < sizeof(fn) / sizeof(fn[0]); i++) n += fn[i] != 0;',
              '    printf("%d %d\\n", n, ehmm_version());', '    return 0;', '}', '']
    src = tmp_path / "abi_link.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi_link"
    libdir = os.path.dirname(ehmm.SO_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:" + os.path.basename(ehmm.SO_PATH), "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == len(names) and int(out[1]) >= 100
