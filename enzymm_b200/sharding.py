"""Multi-GPU sharding: by query structure, no collective on the data path.

Every (template, structure) pair is independent and the only cross-pair state (skip-smaller
flags, completeness) is per structure, so each rank takes a contiguous block of structures, runs
all templates on its own GPU against its own copy of the ~5 MB compiled library, and rank 0
concatenates the per-rank hit lists in input order.  ``torch.distributed`` is used for that one
gather (and for barriers / max-over-ranks timing in ``bench.py``), never inside the search.
"""
from __future__ import annotations

import sys
from typing import List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["shard_bounds", "merge_hits", "gather_hits", "gather_hit_blocks", "gather_hit_buffer", "place_blocks",
           "ChunkQueue"]


def shard_bounds(n_items: int, world_size: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of ``n_items`` for ``rank``; block edges are multiples of
    ``align`` (the synthetic generator's chunk size) except the last."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    units = (n_items + align - 1) // align
    base, extra = divmod(units, world_size)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(lo_u * align, n_items), min(hi_u * align, n_items)


class ChunkQueue:
    """Dynamic hand-out of structure chunks to ranks (for corpora whose structures
    differ widely in cost a static block per rank leaves GPUs idle).  The queue is one counter in
    the process group's key-value store -- host side only, nothing on the data path: every rank
    calls ``next()`` until it returns None and searches chunk ``[lo, hi)`` of the corpus each time.
    Without an initialised process group it simply enumerates the chunks.

    Constructing a queue is collective: every rank creates its queues in the same order, and the
    n-th queue of a job gets its own counter key (``<name>/<n>``), so a second corpus, a second
    ``scan_files`` call or a retry starts from zero instead of from the exhausted counter of the
    previous queue."""

    _created = 0            # queues constructed by this process so far (same on every rank)

    def __init__(self, n_items: int, chunk: int, name: str = "emm_chunk_queue"):
        if chunk <= 0:
            raise ValueError("chunk must be positive")
        self.n_items, self.chunk = int(n_items), int(chunk)
        self.n_chunks = (self.n_items + self.chunk - 1) // self.chunk
        self._local = 0
        self._store = None
        dist = sys.modules.get("torch.distributed")     # not imported -> no process group to ask
        if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._store = dist.distributed_c10d._get_default_store()
        ChunkQueue._created += 1
        self._key = f"{name}/{ChunkQueue._created}"

    def next(self) -> Optional[Tuple[int, int]]:
        if self._store is not None:
            index = int(self._store.add(self._key, 1)) - 1
        else:
            index, self._local = self._local, self._local + 1
        if index >= self.n_chunks:
            return None
        lo = index * self.chunk
        return lo, min(lo + self.chunk, self.n_items)

    def __iter__(self):
        while True:
            span = self.next()
            if span is None:
                return
            yield span


def merge_hits(parts: Sequence[Tuple[int, np.ndarray]]) -> np.ndarray:
    """Concatenate per-shard hit arrays; ``parts`` = (first structure index of the shard, hits with
    shard-local structure indices).  Result is sorted by (structure, template_index)."""
    rebased = []
    for first, hits in parts:
        h = hits.copy()
        h["structure"] += first
        rebased.append(h)
    if not rebased:
        raise ValueError("no shards")
    out = np.concatenate(rebased)
    return out[np.lexsort((out["template_index"], out["structure"]))]


def gather_hits(first: int, hits: np.ndarray, dst: int = 0) -> Optional[np.ndarray]:
    """Gather every rank's (first, hits) on ``dst`` and merge; other ranks get None.
    Works without an initialised process group (single process)."""
    dist = sys.modules.get("torch.distributed")         # not imported -> no process group
    if dist is None or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return merge_hits([(first, hits)])
    payload = (int(first), hits)
    gathered: Optional[List] = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if dist.get_rank() != dst:
        return None
    return merge_hits(gathered)


def place_blocks(blocks: Sequence[Tuple[int, np.ndarray]]) -> np.ndarray:
    """Hit lists of disjoint structure ranges -> one list in input order.  ``blocks`` = (first structure
    of the range, its hits with range-local structure indices, already sorted by (structure,
    template)); ranges are placed by their first structure, so no global sort is needed."""
    blocks = sorted(blocks, key=lambda b: b[0])
    total = sum(len(h) for _, h in blocks)
    if not blocks:
        raise ValueError("no blocks")
    out = np.empty(total, dtype=blocks[0][1].dtype)
    at = 0
    for first, hits in blocks:
        out[at:at + len(hits)] = hits
        out["structure"][at:at + len(hits)] += first
        at += len(hits)
    return out


def gather_hit_blocks(blocks: Sequence[Tuple[int, np.ndarray]], dtype, dst: int = 0, device=None) -> Optional[np.ndarray]:
    """The host-side merge of a run in which every rank searched some chunks of ONE input list
    (``ChunkQueue``): all hit records travel to ``dst`` as raw bytes in one padded ``gather`` (NCCL when
    ``device`` is a CUDA device, else the group's CPU backend), the small (first, count) lists through
    ``all_gather_object``; ``dst`` places every block at its position in input order
    (``place_blocks``).  Other ranks get None.  Single process: just ``place_blocks``."""
    dist = sys.modules.get("torch.distributed")
    if dist is None or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return place_blocks(blocks) if blocks else np.zeros(0, dtype=dtype)
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    meta = [(int(first), int(len(hits))) for first, hits in blocks]
    every = [None] * world
    dist.all_gather_object(every, meta)
    counts = [sum(n for _, n in m) for m in every]
    width = max(max(counts), 1) * np.dtype(dtype).itemsize
    mine = np.concatenate([h for _, h in blocks]) if blocks else np.zeros(0, dtype=dtype)
    send = torch.zeros(width, dtype=torch.uint8, device=device)
    if len(mine):
        send[:mine.nbytes] = torch.from_numpy(mine.view(np.uint8).reshape(-1)).to(send.device)
    recv = [torch.empty(width, dtype=torch.uint8, device=device) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return None
    placed = []
    for r in range(world):
        raw = recv[r][:counts[r] * np.dtype(dtype).itemsize].cpu().numpy().view(dtype)
        at = 0
        for first, n in every[r]:
            placed.append((first, raw[at:at + n]))
            at += n
    return place_blocks(placed) if placed else np.zeros(0, dtype=dtype)


def gather_hit_buffer(buffer, blocks: Sequence[Tuple[int, int, int]], itemsize: int, dst: int = 0, device=None,
                      out=None):
    """The merge of ``gather_hit_blocks`` without host-side copies, for runs whose hit lists are large
    (10^6 structures: ~1.8 GB of records).  ``buffer`` is this rank's ``torch.uint8`` host tensor
    (pinned when CUDA is used) into which every chunk's hit records were downloaded back to back,
    already carrying GLOBAL structure indices; ``blocks`` = (first structure of the chunk, offset in
    records, count) per chunk.  One padded ``gather`` moves the bytes to ``dst``; there every block is
    copied straight from the receive buffers to its place in the merged (pinned) tensor, in input
    order.  Returns that tensor on ``dst`` (``.numpy().view(HIT_DTYPE)`` is the merged hit list),
    None elsewhere.  ``out``: a preallocated (pinned) ``torch.uint8`` tensor on ``dst`` to merge into --
    pinning gigabytes takes longer than moving them.  Single process: the used prefix of ``buffer`` when the blocks are already in
    input order, else a reordered copy."""
    import torch
    dist = sys.modules.get("torch.distributed")
    multi = dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    meta = [(int(f), int(o), int(n)) for f, o, n in blocks]
    if not multi:
        used = sum(n for _, _, n in meta) * itemsize
        if meta == sorted(meta) and all(a[1] + a[2] == b[1] for a, b in zip(meta, meta[1:])):
            return buffer[:used]
        out = torch.empty(used, dtype=torch.uint8)
        at = 0
        for _, off, n in sorted(meta):
            out[at:at + n * itemsize] = buffer[off * itemsize:(off + n) * itemsize]
            at += n * itemsize
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    every = [None] * world
    dist.all_gather_object(every, meta)
    counts = [sum(n for _, _, n in m) for m in every]
    width = max(max(counts), 1) * itemsize
    send = torch.empty(width, dtype=torch.uint8, device=device)
    mine = counts[rank] * itemsize
    send[:mine].copy_(buffer[:mine], non_blocking=True)
    recv = [torch.empty(width, dtype=torch.uint8, device=device) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return None
    total = sum(counts) * itemsize
    if out is not None and out.numel() >= total:
        merged = out
    else:
        merged = torch.empty(max(total, 1), dtype=torch.uint8)
        if device is not None and str(device).startswith("cuda"):
            merged = merged.pin_memory()
    order = sorted((first, r, off, n) for r, m in enumerate(every) for first, off, n in m)
    at = 0
    for first, r, off, n in order:
        nb = n * itemsize
        merged[at:at + nb].copy_(recv[r][off * itemsize:off * itemsize + nb], non_blocking=True)
        at += nb
    if device is not None and str(device).startswith("cuda"):
        torch.cuda.synchronize()
    return merged[:total]
