"""Multi-GPU layer: one process per GPU, `torch.distributed` for the plumbing.

* add_documents / vectorise: units (chunks) are independent -> contiguous partition of the ordered chunk list across
  ranks (keeps the doc <-> embedding order the handlers rely on, tensor_fields_container.py:220-223); NO collective —
  each rank appends its embeddings to its own row-store shard.
* search: the corpus is row-sharded by document; every rank scans its shard for the same query block and produces a
  local top-k; the [nq, k] (doc, row, score) blocks are exchanged ONCE and every rank merges the world_size * k
  candidates per query under the same total order (score desc, doc asc) -> identical result on all ranks.
  Exchange, fastest available first:
    "peer"   — b200_index_search_exchange: ONE kernel stores the packed 16-byte-per-hit block into every peer's
               symmetric buffer over NVLink (CUDA IPC mappings), publishes it with a release flag, waits for the
               peers' blocks and merges them.  No NCCL call on the query path.
    "nccl"   — one all_gather_into_tensor of the packed block + b200_topk_merge_device.
    "host"   — one all_gather of the packed block as a CPU byte tensor + b200_topk_merge (gloo; the CPU tests).
The reference has no counterpart (replicas only: api_validation.py:49-68, s2_inference.py:276-281).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

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


def pack_block(doc: np.ndarray, row: np.ndarray, score: np.ndarray) -> np.ndarray:
    """[nq, k] lists -> the packed block {int32 doc | int32 row | f64 score} the device path exchanges."""
    return np.concatenate([np.ascontiguousarray(doc, np.int32).view(np.uint8).ravel(),
                           np.ascontiguousarray(row, np.int32).view(np.uint8).ravel(),
                           np.ascontiguousarray(score, np.float64).view(np.uint8).ravel()])


def unpack_blocks(blocks: np.ndarray, world: int, nq: int, k: int):
    nk = nq * k
    b = np.ascontiguousarray(blocks, np.uint8).reshape(world, nk * 16)
    doc = b[:, :nk * 4].copy().view(np.int32).reshape(world, nq, k)
    row = b[:, nk * 4:nk * 8].copy().view(np.int32).reshape(world, nq, k)
    score = b[:, nk * 8:].copy().view(np.float64).reshape(world, nq, k)
    return doc, row, score


def allgather_topk(doc: np.ndarray, row: np.ndarray, score: np.ndarray, group=None, device=None):
    """All-gather per-shard lists [nq, k] (doc ids already GLOBAL) as ONE packed block and merge them on the host.
    Works on any backend: pass `device` = the rank's CUDA device for NCCL, leave None for gloo."""
    import torch
    import torch.distributed as dist
    from .engine import topk_merge
    world = dist.get_world_size(group)
    nq, k = doc.shape
    blk = torch.from_numpy(pack_block(doc, row, score))
    if device is not None:
        blk = blk.to(device)
    out = torch.empty(world * blk.numel(), dtype=torch.uint8, device=blk.device)
    dist.all_gather_into_tensor(out, blk, group=group)
    D, R, S = unpack_blocks(out.cpu().numpy(), world, nq, k)
    return topk_merge(D, R, S)


class ShardedRowStore:
    """A row-sharded index: this rank owns documents [doc_base, doc_base + local docs) of the global numbering.
    `device`: the rank's torch CUDA device (None = host exchange, for the gloo tests)."""

    MAX_NQ = 64

    def __init__(self, store, rank: int, world: int, group=None, device=None, exchange: str = "auto", max_k: int = 16):
        self.store = store
        self.rank, self.world, self.group, self.device = rank, world, group, device
        self.doc_base = 0           # global document number of this shard's document 0
        self.max_k = max_k
        self.mode = "host"
        self._ex = None
        self._bufs = None
        self._stream = None
        if device is not None:
            import torch
            self._stream = torch.cuda.Stream(device=device)      # queries in, scan, exchange, block out: one stream
            store.set_stream(self._stream.cuda_stream)
        if device is not None and world > 1:
            self.mode = "nccl"
            if exchange in ("auto", "peer") and world * max_k <= 256:
                try:
                    self._open_peer_exchange()
                    self.mode = "peer"
                except Exception:
                    if exchange == "peer":
                        raise
        elif device is not None:
            self.mode = "single"

    # -- setup ----------------------------------------------------------------------------------------------------
    def _open_peer_exchange(self) -> None:
        import torch
        import torch.distributed as dist
        from .engine import Exchange
        ex = Exchange(self.device.index, self.rank, self.world, max_nq=self.MAX_NQ, max_k=self.max_k)
        mine = torch.frombuffer(bytearray(ex.handle), dtype=torch.uint8).to(self.device)
        allh = torch.empty(self.world * mine.numel(), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        raw = bytes(allh.cpu().numpy().tobytes())
        ok = torch.ones(1, device=self.device)
        try:
            ex.open([raw[i * 64:(i + 1) * 64] for i in range(self.world)])
        except Exception:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)     # every rank takes the same path
        if ok.item() < 1:
            ex.close()
            raise RuntimeError("CUDA IPC peer mapping is not available between all ranks")
        self._ex = ex

    def _device_buffers(self, nq: int, k: int):
        import torch
        key = (nq, k)
        if self._bufs is None or self._bufs[0] != key:
            nk = nq * k
            self._bufs = (key,
                          torch.empty(self.MAX_NQ * self.store.dim, dtype=torch.float32, device=self.device),
                          torch.empty(nk * 16, dtype=torch.uint8, device=self.device),
                          torch.empty(self.world * nk * 16, dtype=torch.uint8, device=self.device),
                          torch.empty(nk * 16, dtype=torch.uint8, device=self.device),
                          torch.empty(nk * 16, dtype=torch.uint8).pin_memory())
        return self._bufs[1:]

    # -- ingest ---------------------------------------------------------------------------------------------------
    def add_local(self, vecs, local_doc_ids: Optional[Sequence[int]], doc_base: int) -> None:
        self.doc_base = int(doc_base)
        self.store.set_doc_offset(self.doc_base)     # the engine returns global document numbers directly
        self.store.add(vecs, local_doc_ids)

    # -- search ---------------------------------------------------------------------------------------------------
    def search_device(self, d_q_ptr: int, nq: int, k: int, d_out_ptr: int, sync: bool = False) -> None:
        """Queries [nq, dim] fp32 on the device -> merged packed block {doc | row | score} at d_out_ptr (nq*k*16 bytes).
        Asynchronous on the store's stream unless sync."""
        import torch.distributed as dist
        nk = nq * k
        _, blk, gathered, _, _ = self._device_buffers(nq, k)
        if self.mode == "peer":
            self.store.search_exchange(self._ex, d_q_ptr, nq, k, blk.data_ptr(), d_out_ptr, d_out_ptr + nk * 4,
                                       d_out_ptr + nk * 8, sync=sync)
        elif self.mode == "nccl":
            import torch
            self.store.search_device(d_q_ptr, nq, k, blk.data_ptr(), blk.data_ptr() + nk * 4, blk.data_ptr() + nk * 8,
                                     sync=False)
            with torch.cuda.stream(self._stream):      # NCCL enqueues on torch's current stream: the store's stream
                dist.all_gather_into_tensor(gathered, blk, group=self.group)
            self.store.merge_shards_device(gathered.data_ptr(), self.world, nq, k, d_out_ptr, d_out_ptr + nk * 4,
                                           d_out_ptr + nk * 8, sync=sync)
        else:
            self.store.search_device(d_q_ptr, nq, k, d_out_ptr, d_out_ptr + nk * 4, d_out_ptr + nk * 8, sync=sync)

    def search(self, queries, k: int):
        """Host queries in -> merged global (doc, row, score) [nq, k] out, identical on every rank.
        One packed exchange per block of <= 64 queries; device modes keep everything but the queries and the final
        block off the host."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if self.mode == "host":
            gdoc, row, score = self.store.search(q, k)
            if self.world == 1:
                return gdoc, row, score
            return allgather_topk(gdoc, row, score, self.group, None)
        import torch
        nq_total = q.shape[0]
        docs, rows, scores = [], [], []
        stream = self._stream
        for lo in range(0, nq_total, self.MAX_NQ):
            nq = min(self.MAX_NQ, nq_total - lo)
            dq, _, _, out, host = self._device_buffers(nq, k)
            with torch.cuda.stream(stream):
                dq[:nq * self.store.dim].copy_(torch.from_numpy(q[lo:lo + nq].ravel()), non_blocking=True)
                self.search_device(dq.data_ptr(), nq, k, out.data_ptr(), sync=False)
                host.copy_(out, non_blocking=True)
            stream.synchronize()
            D, R, S = unpack_blocks(host.numpy(), 1, nq, k)
            docs.append(D[0]); rows.append(R[0]); scores.append(S[0])
        return np.concatenate(docs), np.concatenate(rows), np.concatenate(scores)

    def close(self) -> None:
        if self._ex is not None:
            self._ex.close()
            self._ex = None
