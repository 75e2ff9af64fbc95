"""The C-ABI library loads and exports every symbol include/geneclust_b200.h declares (no GPU needed: no compute
call is made)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "geneclust_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gc_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    from scgeneclust_b200.csrc.build import build
    return build()


def test_header_declares_symbols():
    syms = declared_symbols()
    assert "gc_mi_cc_pairs" in syms and "gc_rsvd_apply" in syms and "gc_kmeans_assign" in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol(built_lib):
    dll = ctypes.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(dll, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    dll.gc_abi_version.restype = ctypes.c_int
    assert dll.gc_abi_version() == 1
    dll.gc_launch_count.restype = ctypes.c_longlong
    assert dll.gc_launch_count() == 0


def test_binding_matches_header(built_lib):
    from scgeneclust_b200._lib import SIGNATURES
    assert sorted(SIGNATURES) == declared_symbols()


def test_binding_arity_matches_header():
    """Every ctypes signature has as many arguments as the C declaration."""
    from scgeneclust_b200._lib import SIGNATURES
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, argtypes) in SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", text, flags=re.S)
        assert m, name
        args = m.group(1).strip()
        n_args = 0 if args in ("", "void") else len(args.split(","))
        assert n_args == len(argtypes), f"{name}: header has {n_args} args, binding {len(argtypes)}"


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from scgeneclust_b200._lib import GeneClustLibError, require_cuda
    with pytest.raises(GeneClustLibError):
        require_cuda()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "scgeneclust_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
