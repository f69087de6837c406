"""World-size-2 run of the sharding path on CPU (gloo): shard by structure, no data-path
collective, rank 0 merges hit lists in input order.  The per-shard search is done
by the oracle here -- the GPU equivalent is tests/test_gpu_parity.py::test_determinism_and_shard_union."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from enzymm_b200.engine import HIT_DTYPE
from enzymm_b200.sharding import gather_hits, merge_hits, shard_bounds
from enzymm_b200.synth import SynthConfig, generate_chunk
from enzymm_b200.templates import load_templates


def test_shard_bounds_cover_everything():
    for n, world, align in ((10, 3, 1), (1000, 8, 256), (5, 8, 1), (0, 2, 1), (1024, 4, 256)):
        spans = [shard_bounds(n, world, r, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b
        for lo, hi in spans[:-1]:
            assert (lo % align == 0 or lo == n) and (hi % align == 0 or hi == n)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _oracle_hits(templates, mols) -> np.ndarray:
    raw = oracle.query_raw(mols, oracle.OracleTemplates(templates), 2.0, 1.5, 1.5, threads=1)
    out = []
    for mi, ti in zip(*np.nonzero(raw["found"])):
        h = np.zeros((), dtype=HIT_DTYPE)
        h["structure"], h["template_index"], h["rmsd"] = mi, ti, raw[mi, ti]["rmsd"]
        h["atoms"] = raw[mi, ti]["atoms"]
        out.append(h)
    return np.array(out, dtype=HIT_DTYPE)


def _inputs():
    templates = list(load_templates(subset="3_residues/results/csa3d_000"))[:30]
    chunk = generate_chunk(5, SynthConfig(n_residues=80, max_motifs=2), templates, 6)
    return templates, [chunk.to_molecule(i) for i in range(6)]


def _worker(rank: int, world: int, port: int, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        templates, mols = _inputs()
        lo, hi = shard_bounds(len(mols), world, rank)
        merged = gather_hits(lo, _oracle_hits(templates, mols[lo:hi]))
        dist.barrier()
        if rank == 0:
            queue.put(merged.tobytes())
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    got = queue.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    templates, mols = _inputs()
    want = merge_hits([(0, _oracle_hits(templates, mols))])
    assert len(want) > 0
    assert got == want.tobytes()


def _queue_worker(rank: int, world: int, port: int, queue):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from enzymm_b200.sharding import ChunkQueue
        templates, mols = _inputs()
        parts = []
        for lo, hi in ChunkQueue(len(mols), chunk=2):          # chunks handed out first come, first served
            parts.append((lo, _oracle_hits(templates, mols[lo:hi])))
        # a second queue in the same job starts from zero again (its own counter key)
        dist.barrier()
        again = [None] * world
        dist.all_gather_object(again, [lo for lo, _ in ChunkQueue(len(mols), chunk=3)])
        assert sorted(lo for part in again for lo in part) == [0, 3], again
        mine = merge_hits(parts) if parts else np.zeros(0, dtype=HIT_DTYPE)
        merged = gather_hits(0, mine)                           # already rebased to corpus indices
        spans = [None] * world
        dist.all_gather_object(spans, [p[0] for p in parts])
        # the block-wise merge of the strong-scaling path: raw bytes in one gather, placed by range
        from enzymm_b200.sharding import gather_hit_blocks
        placed = gather_hit_blocks(parts, HIT_DTYPE)
        # ... and its copy-free form: hits downloaded back to back into one buffer, global indices
        import torch
        from enzymm_b200.sharding import gather_hit_buffer
        room = torch.zeros(64 * HIT_DTYPE.itemsize, dtype=torch.uint8)
        records, blocks, used = room.numpy().view(HIT_DTYPE), [], 0
        for lo, hits in parts:
            records[used:used + len(hits)] = hits
            records["structure"][used:used + len(hits)] += lo
            blocks.append((lo, used, len(hits)))
            used += len(hits)
        buffered = gather_hit_buffer(room, blocks, HIT_DTYPE.itemsize)
        dist.barrier()
        if rank == 0:
            assert placed.tobytes() == merged.tobytes()
            assert buffered.numpy().view(HIT_DTYPE).tobytes() == merged.tobytes()
            queue.put((merged.tobytes(), spans))
        else:
            assert placed is None and buffered is None
    finally:
        dist.destroy_process_group()


def test_dynamic_chunk_queue_covers_every_structure_once():
    from enzymm_b200.sharding import ChunkQueue
    assert list(ChunkQueue(5, 2)) == [(0, 2), (2, 4), (4, 5)] and list(ChunkQueue(0, 4)) == []
    with pytest.raises(ValueError):
        ChunkQueue(4, 0)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.SimpleQueue()
    procs = [ctx.Process(target=_queue_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    got, spans = queue.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(lo for part in spans for lo in part) == [0, 2, 4]      # every chunk exactly once
    templates, mols = _inputs()
    want = merge_hits([(0, _oracle_hits(templates, mols))])
    assert got == want.tobytes()


def test_hit_merges_in_a_single_process():
    """Without a process group the merges are plain host operations: blocks in any order come out in
    input order with global structure indices."""
    import torch
    from enzymm_b200.sharding import gather_hit_blocks, gather_hit_buffer, place_blocks
    def block(first, structures):
        h = np.zeros(len(structures), dtype=HIT_DTYPE)
        h["structure"] = structures
        h["template_index"] = np.arange(len(structures)) + first
        return h
    blocks = [(8, block(8, [0, 0, 3])), (0, block(0, [1, 2])), (4, block(4, [0]))]
    merged = place_blocks(blocks)
    assert merged["structure"].tolist() == [1, 2, 4, 8, 8, 11] and merged["template_index"].tolist() == [0, 1, 4, 8, 9, 10]
    assert gather_hit_blocks(blocks, HIT_DTYPE).tobytes() == merged.tobytes()
    assert len(gather_hit_blocks([], HIT_DTYPE)) == 0
    # buffer form: records already carry global indices; blocks recorded out of input order
    room = torch.zeros(16 * HIT_DTYPE.itemsize, dtype=torch.uint8)
    records = room.numpy().view(HIT_DTYPE)
    meta, used = [], 0
    for first, hits in blocks:
        records[used:used + len(hits)] = hits
        records["structure"][used:used + len(hits)] += first
        meta.append((first, used, len(hits)))
        used += len(hits)
    out = gather_hit_buffer(room, meta, HIT_DTYPE.itemsize).numpy().view(HIT_DTYPE)
    assert out.tobytes() == merged.tobytes()
    in_order = sorted(meta)
    assert gather_hit_buffer(room, [(0, 0, 3)], HIT_DTYPE.itemsize).numel() == 3 * HIT_DTYPE.itemsize
