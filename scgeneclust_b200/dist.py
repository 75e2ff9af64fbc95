"""Gene-shard communicator: one process per GPU, genes partitioned across ranks (BASELINE.json north_star), NCCL over
NVLink for the few small exchanges of the hot path:

  * all-reduce  : per-cell count sums (int64 [C]), per-cell residual sums (float64 [C]), the cells x 64 panel of
                  `A @ Q` (sum over gene shards), 64 x 64 Gram matrices of gene-space panels;
  * all-gather  : gene-space results (G_p x 50 embedding shards), MI ranges (condensed redundancy values).

The reference has no distributed path at all (its parallelism is `multiprocessing.Pool`, tl/information.py:47,83,130).
`torch.distributed` is plumbing: backend "nccl" on the GPUs, "gloo" in the CPU tests of this host-side logic.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous near-equal split of range(total) over `world` ranks; shard sizes are multiples of 64 (the
    kernels' column granularity) except possibly the last one."""
    per = -(-total // world)
    per = -(-per // 64) * 64
    out = []
    for r in range(world):
        lo = min(total, r * per)
        out.append((lo, min(total, lo + per)))
    return out


class GeneShardComm:
    def __init__(self, total_genes: int, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.total_genes = int(total_genes)
        self.bounds = shard_bounds(self.total_genes, self.world)
        if any(hi <= lo for lo, hi in self.bounds):
            # shards are multiples of 64 genes: with fewer than 64 * (world - 1) + 1 genes a rank would hold nothing and
            # its kernels would refuse the empty matrix while the other ranks wait in a collective.  Every rank computes
            # the same bounds, so every rank raises here.
            raise ValueError(f"{self.total_genes} genes cannot be split over {self.world} ranks in shards of 64: "
                             f"use fewer ranks (at most {max(1, -(-self.total_genes // 64))})")
        self.gene_lo, self.gene_hi = self.bounds[self.rank]

    @property
    def local_genes(self) -> int:
        return self.gene_hi - self.gene_lo

    # -- reductions ----------------------------------------------------------------------------------------
    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def absmax_merge(self, v: torch.Tensor) -> torch.Tensor:
        """Element-wise: the signed value of largest magnitude over ranks, lowest rank on ties (rank order = gene
        order, so this is the first-index rule of np.argmax(abs(.)) in svd_flip)."""
        if self.world == 1:
            return v
        allv = [torch.empty_like(v) for _ in range(self.world)]
        dist.all_gather(allv, v.contiguous(), group=self.group)
        stacked = torch.stack(allv)                                  # world x n
        best = torch.argmax(stacked.abs(), dim=0, keepdim=True)      # first maximal index
        return torch.gather(stacked, 0, best)[0].contiguous()

    # -- gathers -------------------------------------------------------------------------------------------
    def all_gather_rows(self, local: torch.Tensor) -> torch.Tensor:
        """Concatenate the per-rank gene-row shards (rank r holds rows bounds[r]) into the full matrix."""
        if self.world == 1:
            return local
        per = max(hi - lo for lo, hi in self.bounds)
        pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad, group=self.group)
        return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, self.bounds)], dim=0).contiguous()

    # -- row-sliced panels (cells x 64 panels that are sums over gene shards) ------------------------------
    def row_chunk(self, n_rows: int) -> int:
        """Rows per rank when an n_rows panel is cut into `world` equal chunks (the last ones zero padded)."""
        return -(-n_rows // self.world)

    def reduce_scatter_rows(self, full_padded: torch.Tensor, out_chunk: torch.Tensor) -> torch.Tensor:
        """out_chunk = this rank's row chunk of the SUM over ranks of `full_padded` (world * chunk rows)."""
        if self.world == 1:
            out_chunk.copy_(full_padded)
        elif dist.get_backend(self.group) == "gloo":          # gloo has no reduce-scatter: CPU-test / one-GPU path
            dist.all_reduce(full_padded, op=dist.ReduceOp.SUM, group=self.group)
            c = out_chunk.shape[0]
            out_chunk.copy_(full_padded[self.rank * c:(self.rank + 1) * c])
        else:
            dist.reduce_scatter_tensor(out_chunk, full_padded, op=dist.ReduceOp.SUM, group=self.group)
        return out_chunk

    def all_gather_chunks(self, chunk: torch.Tensor, out_full: torch.Tensor) -> torch.Tensor:
        """out_full (world * chunk rows) = the per-rank chunks in rank order."""
        if self.world == 1:
            out_full.copy_(chunk)
        else:
            dist.all_gather_into_tensor(out_full, chunk.contiguous(), group=self.group)
        return out_full

    def split_range(self, n: int) -> Tuple[int, int]:
        """This rank's contiguous slice of range(n) (e.g. of the condensed gene-pair index space)."""
        per = -(-n // self.world)
        lo = min(n, self.rank * per)
        return lo, min(n, lo + per)

    def all_gather_ranges(self, local: torch.Tensor, n: int) -> torch.Tensor:
        """Inverse of `split_range`: concatenate the per-rank 1-D slices into the length-n vector."""
        if self.world == 1:
            return local
        per = -(-n // self.world)
        pad = torch.zeros(per, dtype=local.dtype, device=local.device)
        pad[:local.numel()] = local
        parts = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(parts, pad, group=self.group)
        return torch.cat(parts)[:n].contiguous()

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)
