"""B200-native GeneClust gene-grouping hot path: `scGeneClust(adata, version='fast'|'ps', ...)` with the reference's
signature and AnnData in/out contract (scGeneClust/__init__.py:6), running on hand-written sm_100a CUDA kernels
behind the C ABI of libgeneclust_b200.so (include/geneclust_b200.h).  There is no CPU fallback."""
from ._anndata_lite import AnnDataLite, is_anndata
from ._model import scGeneClust

__all__ = ["scGeneClust", "AnnDataLite", "is_anndata"]
__version__ = "0.1.0"
