"""Gene-clustering stage functions, same names as the reference's `scGeneClust.tl` (tl/__init__.py:7-9)."""
from .cluster import cluster_genes
from .information import find_relevant_genes
from .selection import select_from_clusters

__all__ = ["cluster_genes", "find_relevant_genes", "select_from_clusters"]
