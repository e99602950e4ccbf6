"""CPU: the C-ABI library loads and exports every symbol include/gmrm_b200.h declares.
No compute call is made (there is no GPU here); creating an engine must fail loudly."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gmrm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gmrm_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_declared_symbols():
    from gmrm_b200 import api
    lib = api.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gmrm_b200.h but not exported"


def test_no_cpu_fallback():
    import torch
    from gmrm_b200 import api
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.GmrmError, match="no CUDA device|CPU"):
        api.Engine(N=100, Mt=10)


def test_product_does_not_reference_oracle():
    # the product tree must not include, link or import anything under oracle/
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "gmrm_b200")):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp", ".py")) or f == "Makefile":
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r'(#include\s*["<][^">]*oracle|from oracle|import oracle|-loracle|oracle/)', txt):
                    bad.append(os.path.join(base, f))
    assert not bad, bad
