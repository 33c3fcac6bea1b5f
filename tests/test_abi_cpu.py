"""The C-ABI library loads and exports exactly what include/*.h declares (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    return set(re.findall(r"\b((?:hpfw|par_collector|prepare_result|calc_hashprint_result)_\w+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    from hpfw_b200 import _lib
    L = _lib.load()
    decl = _declared("hpfw_b200.h")
    assert decl, "header parse found nothing"
    for name in sorted(decl):
        assert hasattr(L, name), f"{name} declared in include/hpfw_b200.h but not exported"
    assert decl == set(_lib.ABI_SYMBOLS), decl ^ set(_lib.ABI_SYMBOLS)


def test_pyhpfw_abi_symbols_exported():
    path = os.path.join(ROOT, "include", "hpfw_b200_pyhpfw.h")
    if not os.path.exists(path):
        pytest.skip("reference-compatible par_collector_* header not present yet")
    from hpfw_b200 import _lib
    L = _lib.load()
    for name in sorted(_declared("hpfw_b200_pyhpfw.h")):
        assert hasattr(L, name), name


def test_no_cpu_fallback_without_device():
    """On a box without a GPU every entry point must fail loudly with HPFW_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hpfw_b200 import _lib
    L = _lib.load()
    h = C.c_void_p()
    assert L.hpfw_ctx_create(0, C.byref(h)) == _lib.ERR_CUDA
    assert b"no CPU fallback" in L.hpfw_last_error()
    import hpfw_b200
    with pytest.raises(hpfw_b200.HpfwError):
        hpfw_b200.Context(0)


def test_static_helpers():
    from hpfw_b200 import _lib
    L = _lib.load()
    assert L.hpfw_hashprint_words_for_cols(1210) == 1111
    assert L.hpfw_hashprint_words_for_cols(99) == 0
    assert b"sm_100a" in L.hpfw_version()


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (checked textually over the package and headers)."""
    bad = []
    for base in ("hpfw_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if "_build" in dp or "__pycache__" in dp:
                continue
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dp, fn), errors="replace").read()
                    # imports, dlopen targets or include paths; a comment that cites oracle/nsgcq.py as the algorithm
                    # statement is fine
                    if re.search(r"^\s*(import|from)\s+oracle\b|#include\s*[\"<][^\n]*oracle|libhpfw_oracle|libhpfw_ref|"
                                 r"_ref/", txt, flags=re.M):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
