"""Multi-GPU layer: one process per GPU, `torch.distributed` for the plumbing.

* add_documents / vectorise: units (chunks) are independent -> contiguous partition of the ordered chunk list across
  ranks (keeps the doc <-> embedding order the handlers rely on, tensor_fields_container.py:220-223); NO collective —
  each rank appends its embeddings to its own row-store shard.
* search: the corpus is row-sharded by document; every rank scans its shard for the same query block and produces a
  local top-k; ONE all-gather of the [nq, k] (doc, row, score) lists; every rank merges the world_size * k candidates
  per query under the same total order (score desc, doc asc) -> identical result on all ranks.
The reference has no counterpart (replicas only: api_validation.py:49-68, s2_inference.py:276-281).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, order-preserving partition: the first n % world ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def partition(items: Sequence, rank: int, world: int) -> Sequence:
    lo, hi = shard_bounds(len(items), rank, world)
    return items[lo:hi]


def allgather_topk(doc: np.ndarray, row: np.ndarray, score: np.ndarray, group=None, device=None):
    """All-gather per-shard lists [nq, k] (doc ids already GLOBAL) and merge them.  Works on any backend: pass
    `device` = the rank's CUDA device for NCCL, leave None for gloo.  Returns (doc, row, score) [nq, k]."""
    import torch
    import torch.distributed as dist
    from .engine import topk_merge
    world = dist.get_world_size(group)
    td = torch.from_numpy(np.ascontiguousarray(doc, dtype=np.int32))
    tr = torch.from_numpy(np.ascontiguousarray(row, dtype=np.int32))
    ts = torch.from_numpy(np.ascontiguousarray(score, dtype=np.float64))
    if device is not None:
        td, tr, ts = td.to(device), tr.to(device), ts.to(device)
    gd = [torch.empty_like(td) for _ in range(world)]
    gr = [torch.empty_like(tr) for _ in range(world)]
    gs = [torch.empty_like(ts) for _ in range(world)]
    dist.all_gather(gd, td, group=group)
    dist.all_gather(gr, tr, group=group)
    dist.all_gather(gs, ts, group=group)
    D = torch.stack(gd).cpu().numpy()
    R = torch.stack(gr).cpu().numpy()
    S = torch.stack(gs).cpu().numpy()
    return topk_merge(D, R, S)


class ShardedRowStore:
    """A row-sharded index: this rank owns documents [doc_lo, doc_hi) of the global numbering."""

    def __init__(self, store, rank: int, world: int, group=None, device=None):
        self.store = store
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.doc_base = 0           # global document number of this shard's document 0

    def add_local(self, vecs, local_doc_ids: Optional[Sequence[int]], doc_base: int) -> None:
        self.doc_base = int(doc_base)
        self.store.set_doc_offset(self.doc_base)     # the engine returns global document numbers directly
        self.store.add(vecs, local_doc_ids)

    def search(self, queries, k: int):
        gdoc, row, score = self.store.search(queries, k)
        if self.world == 1:
            return gdoc, row, score
        return allgather_topk(gdoc, row, score, self.group, self.device)
