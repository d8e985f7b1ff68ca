"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, exports every symbol that
include/kgl_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib_path():
    from kgl_gene_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kgl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kgl_b200_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib_path):
    lib = C.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/kgl_b200.h but not exported"


def test_python_binding_lists_the_same_symbols():
    from kgl_gene_b200 import capi
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_struct_layout_matches_header():
    from kgl_gene_b200 import capi
    assert capi.RESULT_DTYPE.itemsize == 80          # 5 x (uint64 + double)
    assert C.sizeof(capi.InbreedOptions) == 48         # pointer, int32 (+4), double, 5 x int32 (+4)


def test_no_cpu_fallback_without_gpu(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from kgl_gene_b200.capi import KglB200, KglError
    with pytest.raises(KglError, match="no CPU fallback|no CUDA device"):
        KglB200(0)


def test_product_does_not_reference_the_oracle():
    """The oracle is test infrastructure: nothing under kgl_gene_b200/ may import, link or name it."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "kgl_gene_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                s = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"oracle_py|kgl_oracle|libkgl_oracle|oracle/_ref", s):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
