"""ctypes binding of libgeneclust_b200.so (the C ABI declared in include/geneclust_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
Device pointers are taken from torch tensors (`tensor.data_ptr()`); torch is plumbing only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# GENECLUST_B200_LIB: developer override (kernel-variant timing, scripts/pass_times.py); the in-tree build is the default
LIB_PATH = os.environ.get("GENECLUST_B200_LIB") or os.path.join(_HERE, "libgeneclust_b200.so")

_p = C.c_void_p
_i32, _i64, _u64, _f32, _f64 = C.c_int32, C.c_int64, C.c_uint64, C.c_float, C.c_double


class ResidualMatrix(C.Structure):
    """`gc_residual_matrix` (include/geneclust_b200.h)."""
    _fields_ = [("counts", _p), ("C", _i64), ("G", _i64), ("ld", _i64),
                ("row_sum_f", _p), ("col_sum_f", _p), ("total", _f32), ("theta", _f32), ("clip", _f32),
                ("row_shift", _p), ("col_shift", _p), ("upper_clip_inactive", _i32)]


# name -> (restype, argtypes).  Must list every function declared in include/geneclust_b200.h
# (tests/test_abi.py checks both directions).
SIGNATURES = {
    "gc_abi_version": (C.c_int, []),
    "gc_last_error": (C.c_char_p, []),
    "gc_launch_count": (C.c_longlong, []),
    "gc_profile_enable": (C.c_int, [C.c_int]),
    "gc_profile_collect": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "gc_fp64_probe": (C.c_int, [C.POINTER(C.c_double), _p, _p]),
    "gc_csr_to_counts": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _p, _i64, C.c_int, _p, _p]),
    "gc_dense_to_counts": (C.c_int, [_p, _i64, _i64, _i64, _i64, _p, _i64, _p, _p]),
    "gc_count_sums": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "gc_count_stats": (C.c_int, [_p, _i64, _p, _p, _i64, C.c_double, _p, _p, _p, _p, _p]),
    "gc_pearson_residuals": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _f32, _f32, _f32, _p, _i64, _p]),
    "gc_residual_sums": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _f32, _f32, _f32, _p, _p, _p]),
    "gc_rsvd_scratch_bytes": (_i64, [_i64, _i64]),
    "gc_rsvd_apply": (C.c_int, [C.POINTER(ResidualMatrix), C.c_int, _p, _p, C.c_int, _p, _p]),
    "gc_rsvd_apply_ss": (C.c_int, [C.POINTER(ResidualMatrix), C.c_int, _p, _p, C.c_int, _p, _p]),
    "gc_rsvd_apply_simt": (C.c_int, [C.POINTER(ResidualMatrix), C.c_int, _p, _p, C.c_int, _p, _p]),
    "gc_mt19937_normal_scratch_bytes": (_i64, [_i64, C.c_int]),
    "gc_mt19937_normal_panel": (C.c_int, [_p, C.c_int, _i64, C.c_int, _p, C.c_int, _p, _p, _p]),
    "gc_gram_scratch_bytes": (_i64, [_i64]),
    "gc_panel_gram": (C.c_int, [_p, _i64, _p, _p, _p]),
    "gc_panel_matmul": (C.c_int, [_p, _i64, _p, _p, _p]),
    "gc_panel_colabsmax": (C.c_int, [_p, _i64, _p, _p, _p, _p]),
    "gc_chol_inv": (C.c_int, [_p, C.c_int, _p, _p, _p]),
    "gc_sym_eig": (C.c_int, [_p, C.c_int, _p, _p, _p, _p]),
    "gc_scale_cols": (C.c_int, [_p, _p, _p, C.c_int, _p]),
    "gc_kmeans_assign": (C.c_int, [_p, _i64, _p, _i64, _p, C.c_int, C.c_int, _p, _p, _p]),
    "gc_kmeans_minibatch_update": (C.c_int, [_p, _i64, _p, _i64, _p, _p, _p, _p, C.c_int, C.c_int, _p]),
    "gc_sqdist_to_rows": (C.c_int, [_p, _i64, _i64, C.c_int, _p, C.c_int, _p, _p]),
    "gc_kmeanspp_scratch_bytes": (_i64, [_i64, C.c_int]),
    "gc_kmeanspp_init": (C.c_int, [_p, _i64, _i64, C.c_int, C.c_int, _i32, _p, C.c_int, _p, _p, _p, _p]),
    "gc_searchsorted_right": (C.c_int, [_p, _i64, _p, C.c_int, _p, _p]),
    "gc_sum_f32": (C.c_int, [_p, _i64, _p, _p]),
    "gc_kmeans_minibatch_step": (C.c_int, [_p, _i64, _i64, C.c_int, C.c_int, _p, _p, C.c_int, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gc_kmeans_reassign": (C.c_int, [_p, _i64, C.c_int, _p, _p, _p, C.c_int, _f32, _p, _p, _p]),
    "gc_paired_dist": (C.c_int, [_p, _i64, _i64, C.c_int, _p, _p, _p, _p]),
    "gc_cluster_extremes": (C.c_int, [_p, _p, _i64, C.c_int, _p, _p, _p]),
    "gc_closeness": (C.c_int, [_p, _p, _i64, C.c_int, _p, _p, _p]),
    "gc_binomial_deviance_scratch_bytes": (_i64, [_i64, C.c_int]),
    "gc_binomial_deviance": (C.c_int, [_p, _i64, C.c_int, _p, _p, _p]),
    "gc_sum_axis0_f32": (C.c_int, [_p, _i64, C.c_int, _p, _p, _p, _p]),
    "gc_mi_prepare": (C.c_int, [_p, _i64, _p, _p, C.c_int, _i64, _i64, _p, _p, _p, _p, _p]),
    "gc_mi_cc_pairs": (C.c_int, [_p, _p, _i64, C.c_int, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p]),
    "gc_mi_cc_pairs_scratch_bytes": (_i64, [_i64, C.c_int]),
    "gc_mi_cc_pairs_sorted": (C.c_int, [_p, _p, _i64, C.c_int, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "gc_mi_cc_problems": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _p]),
    "gc_mi_cd": (C.c_int, [_p, _i64, _i64, C.c_int, _p, _p, C.c_int, _f64, _p, _p, _p, _p]),
    "gc_mst_scratch_bytes": (_i64, [_i64]),
    "gc_mst_boruvka": (C.c_int, [_p, _i64, _p, _p, _p, _p]),
    "gc_mst_boruvka_edges": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "gc_comembership_counts": (C.c_int, [_p, _p, _i64, C.c_int, _f64, _p, _p, _p]),
    "gc_host_hdbscan_tree": (C.c_int, [_p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "gc_host_iforest_depths": (C.c_int, [_p, C.c_int, C.c_uint32, C.c_int, _p, _p]),
    "gc_synth_counts": (C.c_int, [_u64, _i64, _i64, _i64, _i64, _p, _p, C.c_int, _p, _p, C.c_int, _p, C.c_int,
                                  C.c_int, _p, _i64, _p]),
}

_NO_STATUS = {"gc_abi_version", "gc_last_error", "gc_launch_count", "gc_rsvd_scratch_bytes", "gc_gram_scratch_bytes",
              "gc_mt19937_normal_scratch_bytes", "gc_binomial_deviance_scratch_bytes", "gc_mi_cc_pairs_scratch_bytes",
              "gc_kmeanspp_scratch_bytes", "gc_mst_scratch_bytes"}


class GeneClustLibError(RuntimeError):
    pass


class _Lib:
    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise GeneClustLibError(
                f"{path} not found: build it with `python -m scgeneclust_b200.csrc.build` "
                "(there is no CPU fallback for the GeneClust hot path)")
        self._dll = C.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(self._dll, name)
            fn.restype, fn.argtypes = restype, argtypes
            setattr(self, "_raw_" + name, fn)
            if name in _NO_STATUS:
                setattr(self, name[3:], fn)
            else:
                setattr(self, name[3:], self._checked(name, fn))
        if self.abi_version() != 1:
            raise GeneClustLibError(f"ABI version mismatch: library reports {self.abi_version()}, binding expects 1")

    def _checked(self, name, fn):
        def call(*args):
            rc = fn(*args)
            if rc != 0:
                raise GeneClustLibError(f"{name} failed: {self._dll.gc_last_error().decode(errors='replace')}")
        call.__name__ = name
        return call


_LIB: Optional[_Lib] = None


def lib() -> _Lib:
    """The loaded library (loaded on first use)."""
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL).  The tensor must be contiguous."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("tensor passed to the C ABI must be contiguous")
    return t.data_ptr()


def stream_handle() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise GeneClustLibError("a CUDA device (B200, sm_100a) is required: the GeneClust hot path has no CPU fallback")
    lib()
