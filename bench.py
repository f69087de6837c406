#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: query structures/sec vs the full shipped M-CSA
library on synthetic AlphaFold-scale structures with planted motifs.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port of Jess, all host threads)

A *step* is one pass of the hot path (upload -> prepare -> search/superpose/filter -> hits) over
one batch of ``--structures`` synthetic 400-residue structures per GPU (BASELINE config 2: 10 000
structures on 1 x B200; weak scaling for N > 1: every rank searches its own 10 000).  The batch
(~1 GB of SoA input per GPU) is far larger than the 126 MB L2, so consecutive steps cannot reuse
cached input.

Printed JSON (rank 0): ``value`` = whole-job structures/s with inputs resident in HBM;
``e2e`` = the same through the C-ABI with pinned HOST buffers, host<->device copies inside the
timed region; ``roofline`` = search kernel against the measured HBM peak (20 B/atom +
160 B/hit algorithmic bytes) -- the kernel is issue/latency bound, so that fraction is tiny by
construction and the issue-side numbers are reported next to it; ``cpu_baseline`` = the oracle
(a CPU port of the reference's Jess path; PyJess itself is not installable offline) on the box's
host cores over a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DEFAULT_DIST = {3: 0.9, 4: 1.7, 5: 2.0, 6: 2.0, 7: 2.0, 8: 2.0}     # Matcher._DEFAULT_JESS_PARAMS
METRIC = "query structures/sec vs full M-CSA library"
UNIT = "structures/s"

_TEMPLATES = None


def active_templates():
    global _TEMPLATES
    if _TEMPLATES is None:
        from enzymm_b200.templates import load_templates
        _TEMPLATES = [t for t in load_templates() if t.effective_size >= 3]
    return _TEMPLATES


def _gen(job):
    from enzymm_b200.synth import SynthConfig, generate_chunk
    chunk_index, count, residues, chains = job
    return generate_chunk(chunk_index, SynthConfig(n_residues=residues, n_chains=chains), active_templates(), count)


def make_workload(rank: int, structures: int, residues: int, chains: int, workers: int):
    """Structures of this rank (weak scaling: rank r owns chunks [r*C, (r+1)*C))."""
    from enzymm_b200.synth import CHUNK, SynthChunk
    per_rank_chunks = (structures + CHUNK - 1) // CHUNK
    jobs = []
    left = structures
    for c in range(per_rank_chunks):
        n = min(CHUNK, left)
        jobs.append((rank * per_rank_chunks + c, n, residues, chains))
        left -= n
    active_templates()
    if workers > 1 and len(jobs) > 1:
        with mp.get_context("fork").Pool(min(workers, len(jobs))) as pool:
            parts = pool.map(_gen, jobs)
    else:
        parts = [_gen(j) for j in jobs]
    atom_off = [np.zeros(1, dtype=np.int64)]
    planted = []
    base_s = base_a = 0
    for p in parts:
        atom_off.append(p.atom_off[1:] + base_a)
        planted.extend((s + base_s, t) for s, t in p.planted)
        base_s += p.n_structures
        base_a += p.n_atoms
    cat = lambda name: np.concatenate([getattr(p, name) for p in parts])
    return SynthChunk(np.concatenate(atom_off), cat("xyz"), cat("kind"), cat("residue"), cat("resnum"),
                      cat("chain"), cat("bfactor"), planted, first_index=jobs[0][0] * CHUNK)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed regions."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, cells in self.rows:
            if len(cells) < 7 or not any(a <= ts <= b for a, b in windows):
                continue
            try:
                sm.append(float(cells[0]))
                mx.append(float(cells[1]))
            except ValueError:
                continue
            for name, cell in zip(names, cells[3:7]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_stamp() -> str:
    """SHA-256 over the sources and build flags of the search kernel; ``profiles/roofline_traffic.json``
    carries the stamp of the build its ncu counters were captured from (``tools/roofline_stamp.py``)."""
    import hashlib
    h = hashlib.sha256()
    csrc = ROOT / "enzymm_b200" / "csrc"
    for name in ("emm_search.cu", "emm_device.cuh", "Makefile"):
        h.update((csrc / name).read_bytes())
    return h.hexdigest()


def committed_counters():
    """ncu counters of the search kernel from the committed ``--set full`` capture, or (None, why)
    when the file is missing or was captured from another build of the kernel (stale)."""
    path = ROOT / "profiles" / "roofline_traffic.json"
    try:
        data = json.loads(path.read_text())
    except Exception:
        return None, "profiles/roofline_traffic.json missing"
    if data.get("kernel_sha256") != kernel_stamp():
        return None, "stale: profiles/roofline_traffic.json was captured from another build of emm_search.cu"
    return data, None


def committed_traffic_bytes(n_structures: int):
    """dram__bytes_read+write of the search kernel per launch, from the committed ncu --set full
    capture (taken at a smaller batch; scaled per structure to this launch's batch)."""
    data, _ = committed_counters()
    return None if data is None else int(data["search_kernel_dram_bytes_per_structure"] * n_structures)


def committed_issue_figure(pairs_per_launch: float, kernel_ms: float, sm_mhz):
    """Issue-slot view of the search kernel (it is instruction-bound, not HBM-bound): warp
    instructions per (template, structure) pair from the committed ncu capture x the pairs of this
    launch / the kernel time measured here, against one warp instruction per SM sub-partition per
    clock (SMs x 4 x SM clock).  Extra information next to the HBM roofline the contract asks for."""
    data, why = committed_counters()
    if data is None:
        return {"bound": "issue", "achieved": None, "frac": None, "stale": True, "note": why}
    try:
        per_pair = float(data["search_kernel_warp_instructions_per_pair"])
        import torch
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        mhz = float(sm_mhz or 1965.0)
        achieved = per_pair * pairs_per_launch / (kernel_ms / 1000.0) / 1e9
        peak = sms * 4 * mhz * 1e6 / 1e9
        return {"bound": "issue", "achieved": achieved, "peak": peak, "unit": "G warp-inst/s", "frac": achieved / peak,
                "warp_inst_per_pair": per_pair, "stale": False,
                "source": "profiles/roofline_traffic.json (ncu smsp__inst_executed.sum, same kernel build)"}
    except Exception:
        return None


def cpu_oracle_rate(workload, n_structures: int, threads: int):
    """Oracle (CPU port) over the first ``n_structures`` structures with ``threads`` threads."""
    import oracle
    templates = active_templates()
    dist = np.asarray([DEFAULT_DIST[min(t.effective_size, 8)] for t in templates])
    mols = [workload.to_molecule(i) for i in range(n_structures)]
    ot = oracle.OracleTemplates(templates)
    t0 = time.perf_counter()
    raw = oracle.query_raw(mols, ot, 2.0, dist, dist, max_candidates=10000, ignore_chain=True, threads=threads)
    dt = time.perf_counter() - t0
    cpu_oracle_rate.evals_per_structure = float(raw["dist_evals"].sum()) / max(n_structures, 1)
    cpu_oracle_rate.raw = raw
    cpu_oracle_rate.molecules = mols
    return n_structures / dt, dt, int(raw["found"].sum())


def parity_against_oracle(hits, raw, molecules, templates):
    """Outside the timed region: the GPU hits of the structures the CPU baseline just searched,
    compared with the oracle's -- hit set, matched atoms in template order, RMSD bit for bit,
    complete-assignment counts, orientation (1e-4) and the filter verdict of every hit."""
    import oracle
    from enzymm_b200.engine import HIT_PASS
    n = raw.shape[0]
    mine = hits[hits["structure"] < n]
    want = {(int(s), int(t)) for s, t in zip(*np.nonzero(raw["found"]))}
    got = {}
    for h in mine:
        got[(int(h["structure"]), int(h["template_index"]))] = h
    mismatches = len(want ^ set(got))
    worst_orient = 0.0
    for key in want & set(got):
        s, t = key
        h, r = got[key], raw[s, t]
        m = int(h["n_atoms"])
        ok = (m == len(templates[t]) and h["atoms"][:m].tolist() == r["atoms"][:m].tolist()
              and float(h["rmsd"]) == float(r["rmsd"]) and int(h["n_complete"]) == int(r["n_complete"]))
        if ok:
            moved = (molecules[s].xyz[r["atoms"][:m]] - r["qbar"]) @ r["rot"].reshape(3, 3).T + r["tbar"]
            orient = oracle.orientation(templates[t], moved)
            worst_orient = max(worst_orient, abs(orient - float(h["orientation"])))
            dist = DEFAULT_DIST[min(templates[t].effective_size, 8)]
            verdict = oracle.predicted_correct(templates[t].effective_size, dist, float(r["rmsd"]), orient)
            ok = abs(orient - float(h["orientation"])) <= 1e-4 and verdict == bool(int(h["flags"]) & HIT_PASS)
        mismatches += 0 if ok else 1
    return {"structures": int(n), "pairs": int(n) * len(templates), "hits_oracle": len(want), "hits_gpu": len(got),
            "mismatches": int(mismatches), "max_orientation_diff": worst_orient,
            "checked": "hit set, atoms, rmsd ==, n_complete ==, orientation <= 1e-4, filter verdict"}


def api_legs(workload, n_files: int, chunk: int = 1024):
    """What a user of the Python-facing API gets, from PDB FILES on local disk (page cache) to results,
    on one GPU -- three entry points, each timed end to end on its second call (the first one allocates
    the device sessions):
      files_to_hits     Matcher.scan_files(paths): native ingest -> GPU -> emm_hit records
      files_to_matches  load_molecules(paths) + Matcher.run(molecules): Molecule objects -> {Molecule: [Match]}
                        (the reference's own call sequence, _cli.py:217-248)
      files_to_tsv      Matcher.scan_to_tsv(paths, file): the reference's results table (_cli.py:270-316)
    and, from the same files written once as a packed corpus (packing.write_corpus):
      corpus_to_hits / corpus_to_tsv   the same two calls on the ``.emmpack`` file: no text is parsed
    """
    import io
    import shutil
    import tempfile
    from enzymm_b200 import jess_run
    distinct = min(n_files, workload.n_structures, 1024)
    texts = [workload.to_pdb(i).encode() for i in range(distinct)]
    need = 2 * sum(len(texts[i % distinct]) for i in range(n_files))
    where = None                                  # RAM-backed if there is room, else the default temp directory
    try:
        if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need:
            where = "/dev/shm"
    except OSError:
        pass
    root = tempfile.mkdtemp(prefix="emm_bench_", dir=where)
    try:
        paths = []
        for i in range(n_files):
            path = os.path.join(root, f"s{i:06d}.pdb")
            with open(path, "wb") as handle:
                handle.write(texts[i % distinct])
            paths.append(path)
        size = sum(len(texts[i % distinct]) for i in range(n_files))
        matcher = jess_run.Matcher(templates=active_templates())
        out = {"files": n_files, "distinct_structures": distinct, "pdb_text_bytes": size, "chunk": chunk,
               "host_threads": host_threads(), "unit": "files/s",
               "note": "second call of each entry point; files in the page cache; one GPU"}
        hits = rows = 0
        for rep in range(2):
            t0 = time.perf_counter()
            hits = sum(len(records) for _, _, records in matcher.scan_files(paths, chunk_size=chunk))
            out["files_to_hits"] = n_files / (time.perf_counter() - t0)
        sample = paths[:min(n_files, 4096)]
        for rep in range(2):
            t0 = time.perf_counter()
            molecules = jess_run.load_molecules(sample)
            t1 = time.perf_counter()
            matches = matcher.run(molecules)
            t2 = time.perf_counter()
            out["files_to_matches"] = len(sample) / (t2 - t0)
            out["files_to_matches_parts"] = {"load_molecules": len(sample) / (t1 - t0), "run": len(sample) / (t2 - t1),
                                             "files": len(sample), "matches": sum(len(v) for v in matches.values())}
            del molecules, matches
        for rep in range(2):
            sink = io.BytesIO()
            t0 = time.perf_counter()
            rows = matcher.scan_to_tsv(paths, sink, chunk_size=chunk)
            out["files_to_tsv"] = n_files / (time.perf_counter() - t0)
            out["tsv_rows"], out["tsv_bytes"] = rows, sink.tell()
        out["hits"] = hits
        # the same files as a packed corpus (packing.write_corpus: parsed once, screened from the mapped columns);
        # an extra of the extra -- a failure here costs only these keys
        try:
            from enzymm_b200.packing import write_corpus
            table_from_text = sink.getvalue()
            corpus = os.path.join(root, "all.emmpack")
            t0 = time.perf_counter()
            write_corpus(paths, corpus)
            out["corpus_write"] = n_files / (time.perf_counter() - t0)
            out["corpus_bytes"] = os.path.getsize(corpus)
            corpus_hits = 0
            for rep in range(2):                  # the corpus four times over: enough chunks for the pipeline to fill
                t0 = time.perf_counter()
                corpus_hits = sum(len(records) for _, _, records in matcher.scan_files([corpus] * 4, chunk_size=4096))
                out["corpus_to_hits"] = 4 * n_files / (time.perf_counter() - t0)
            out["corpus_to_hits_structures"] = 4 * n_files
            for rep in range(2):
                sink = io.BytesIO()
                t0 = time.perf_counter()
                matcher.scan_to_tsv([corpus], sink, chunk_size=chunk)
                out["corpus_to_tsv"] = n_files / (time.perf_counter() - t0)
            out["corpus_equal"] = {"hits": corpus_hits == 4 * hits, "table": sink.getvalue() == table_from_text}
        except Exception as exc:              # noqa: BLE001
            out["corpus_error"] = f"{type(exc).__name__}: {exc}"[:300]
        matcher.close()
        return out
    finally:
        shutil.rmtree(root, ignore_errors=True)


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def base_line(args, world):
    return {
        "metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
        "config": {
            "workload": f"{args.structures} synthetic {args.residues}-residue structures per GPU "
                        f"(planted M-CSA motifs, seed 20230210) x 6780 active templates of "
                        f"jess_templates_20230210, default --jess thresholds (BASELINE config 2)",
            "structures_per_gpu": args.structures, "templates": 6780, "parallelism": f"shard-by-structure x{world}",
            "l2": "inputs (~1 GB/GPU) exceed the 126 MB L2; no flush needed between steps",
        },
    }


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference's Jess path on all host threads (rank 0 only)."""
    if rank != 0:
        return
    threads = host_threads()
    sample = args.cpu_sample or max(16, min(4 * threads, 256))
    workload = make_workload(0, sample, args.residues, args.chains, threads)
    for _ in range(args.warmup):
        cpu_oracle_rate(workload, min(sample, threads), threads)
    times = []
    hits = 0
    for _ in range(args.steps):
        rate, dt, hits = cpu_oracle_rate(workload, sample, threads)
        times.append(dt)
    total = sum(times)
    value = sample * args.steps / total
    line = base_line(args, world)
    line.update({
        "impl": "reference", "value": value, "ms_per_step": 1000.0 * total / args.steps, "dtype": "f64",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} structures x 6780 templates per step (CPU restatement of Jess, not PyJess)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "hits_per_step": hits,
    })
    line["config"]["sample_per_step"] = sample
    print(json.dumps(line), flush=True)


def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    workers = max(1, host_threads() // max(world, 1))
    t_gen = time.perf_counter()
    workload = make_workload(rank, args.structures, args.residues, args.chains, workers)
    t_gen = time.perf_counter() - t_gen

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from enzymm_b200.engine import Engine, HIT_PASS, PackedBatch, Session
    from enzymm_b200.library import CompiledLibrary

    templates = active_templates()
    dists = [DEFAULT_DIST[min(t.effective_size, 8)] for t in templates]
    engine = Engine(CompiledLibrary(templates, 2.0, dists, dists), device=local_rank)
    packed = workload.to_packed(engine.compiled)

    def pinned(a):
        return None if a is None else torch.from_numpy(a).pin_memory().numpy()

    host = PackedBatch(pinned(packed.atom_off), pinned(packed.xyz), pinned(packed.klass), pinned(packed.residue),
                       pinned(packed.bfactor), None, None)
    session = engine.session_for(host.n_atoms, host.n_structures, hit_capacity=64 * host.n_structures)
    # everything below runs on explicit streams: the legacy default stream would serialise the two
    # lanes of the end-to-end pipeline (blocking streams synchronise with it)
    torch.cuda.set_stream(torch.cuda.Stream())
    stream = torch.cuda.current_stream().cuda_stream
    run_kwargs = dict(max_candidates=10000, ignore_chain=True, reset=True, force_prepare=True, stream=stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    per_rank = {}

    def max_over_ranks(ms: float, label: str = "") -> float:
        if world == 1:
            per_rank[label] = [ms]
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        per_rank[label] = [float(x.item()) for x in every]
        return max(per_rank[label])

    hits = None
    for _ in range(args.warmup):
        session.upload(host, stream=stream)
        session.run(**run_kwargs)
        hits = session.download(stream=stream)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    windows = []

    # ---- resident path: inputs already in HBM, K x (prepare + search) -------------------------------
    session.upload(host, stream=stream)
    session.clear_timings()
    barrier()
    w0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        session.run(**run_kwargs)
    e1.record()
    barrier()
    windows.append((w0, time.time()))
    resident_ms = max_over_ranks(e0.elapsed_time(e1), "resident")
    search_ms = session.kernel_ms("search")
    prepare_ms = session.kernel_ms("prepare")
    launches = session.last_launches * args.steps
    hits = session.download(stream=stream)

    # ---- end to end: pinned host buffers -> hits on the host, every step ----------------------------
    # Two sessions on two streams: while step i is being searched, step i+1's batch is already
    # crossing PCIe (every step still pays its own H2D copy, prepare, search and D2H of its hits).
    second = Session(engine.device_library, host.n_atoms, host.n_structures, session.hit_capacity)
    side = torch.cuda.Stream()
    lanes = [(session, stream), (second, side.cuda_stream)]
    e2e_kwargs = dict(run_kwargs)

    def submit(i):
        sess, st = lanes[i % 2]
        e2e_kwargs["stream"] = st
        sess.upload(host, stream=st)
        sess.run(**e2e_kwargs)

    for i in range(2):                         # warm the second lane
        submit(i)
        lanes[i % 2][0].download(stream=lanes[i % 2][1])
    barrier()
    w0 = time.time()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    d2h = 0
    submit(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            submit(i + 1)
        sess, st = lanes[i % 2]
        hits = sess.download(stream=st)
        d2h = 16 + hits.nbytes
    torch.cuda.current_stream().wait_stream(side)
    f1.record()
    barrier()
    windows.append((w0, time.time()))
    e2e_ms = max_over_ranks(f0.elapsed_time(f1), "e2e")
    clocks = sampler.stop(windows) if sampler else None
    second.close()

    # ---- sanity outside the timed regions: planted motifs are being found ---------------------------
    found = set(zip(hits["structure"].tolist(), hits["template_index"].tolist()))
    recovered = sum(1 for p in workload.planted if p in found)
    n_pass = int(((hits["flags"] & HIT_PASS) != 0).sum())

    total_structures = world * host.n_structures
    value = total_structures * args.steps / (resident_ms / 1000.0)
    e2e_value = total_structures * args.steps / (e2e_ms / 1000.0)

    if rank != 0:
        return
    peak, peak_src = measured_peak_gbs()
    avg_search_ms = sum(search_ms) / max(len(search_ms), 1)
    alg_bytes = 20 * host.n_atoms + 160 * len(hits)             # 20 B/atom + 160 B/hit
    achieved = alg_bytes / (avg_search_ms / 1000.0) / 1e9 if avg_search_ms else 0.0
    line = base_line(args, world)
    line.update({
        "impl": "b200", "value": value, "ms_per_step": resident_ms / args.steps, "dtype": "f32 (+f64 guard band and superposition)",
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(host.nbytes()),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": committed_traffic_bytes(host.n_structures),
                     "peak_source": peak_src, "kernel": "emm_search_kernel",
                     "kernel_ms_avg": avg_search_ms, "prepare_kernel_ms_avg": sum(prepare_ms) / max(len(prepare_ms), 1),
                     "algorithmic_bytes_per_launch": int(alg_bytes),
                     "note": "gather/compare-bound search: issue and latency bind long before HBM does"},
        "roofline_issue": committed_issue_figure(host.n_structures * len(templates), avg_search_ms,
                                                  (clocks or {}).get("sm_mhz")) if avg_search_ms else None,
        "pairs_per_s": value * len(templates),
        "hits_per_step": int(len(hits)), "hits_passing_filter": n_pass,
        "planted_recovered": f"{recovered}/{len(workload.planted)}",
        "atoms_per_structure": host.n_atoms / max(host.n_structures, 1),
        "clocks": clocks, "generation_s": t_gen,
        "per_rank_ms": {k: [round(v / args.steps, 2) for v in vals] for k, vals in per_rank.items()},
    })
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        probe_rate, _, _ = cpu_oracle_rate(workload, min(threads, host.n_structures), threads)
        sample = int(min(max(threads, probe_rate * 15.0), 512, host.n_structures))
        rate, dt, _ = cpu_oracle_rate(workload, sample, threads)
        # SURVEY 8(d) issue metric: pairwise-distance evaluations of the canonical (oracle) search x 9
        # thread-instructions / (SMs x 4 schedulers x 32 lanes x SM clock)
        evals = getattr(cpu_oracle_rate, "evals_per_structure", None)
        if evals and line.get("roofline_issue"):
            # the peak needs no ncu capture: SMs x 4 schedulers x SM clock (x 32 lanes)
            sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
            lanes_per_s = sms * 4 * float((clocks or {}).get("sm_mhz") or 1965.0) * 1e6 * 32
            line["roofline_issue"]["canonical_evals_per_structure"] = evals
            line["roofline_issue"]["canonical_frac"] = evals * 9.0 * value / lanes_per_s
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"first {sample} structures of the same batch x 6780 templates, {dt:.1f} s "
                                          "(CPU restatement of Jess, not PyJess)"}
        # the oracle's hits of that sample vs the GPU's hits of the same structures
        line["parity"] = parity_against_oracle(hits, cpu_oracle_rate.raw, cpu_oracle_rate.molecules, templates)
    if world == 1 and args.api_files > 0:
        engine.close()              # the API legs bring their own Matcher (own device library and sessions)
        try:
            line["e2e_api"] = api_legs(workload, args.api_files)
        except Exception as exc:              # noqa: BLE001 -- an extra of the line, never a reason to lose it
            line["e2e_api"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    print(json.dumps(line), flush=True)
    engine.close()
    if world > 1:
        dist.destroy_process_group()
    if line.get("parity", {}).get("mismatches"):
        sys.exit(f"PARITY FAILURE: {line['parity']}")


def run_strong(args, rank, local_rank, world):
    """BASELINE config 3: ONE list of ``--total`` structures searched by N GPUs with the host merge
    inside the timed region.  Ranks pull chunks of the list from one counter (``sharding.ChunkQueue``,
    a key in the process group's store: dynamic hand-out, no data-path collective), every chunk is
    uploaded from pinned host memory, prepared, searched and its hits read back (two sessions on two
    streams per rank, as the e2e leg of the weak-scaling bench), and at the end all hit records travel
    to rank 0 in one NCCL gather and are placed in input order (``sharding.gather_hit_buffer``).
    Wall clock from a barrier before the first chunk to the merged list on rank 0.

    Host memory: a 10^6-structure list is ~107 GB of SoA columns, so structure i of the list is
    structure ``i mod D`` of D distinct synthetic structures (``--distinct``, default 16 384 = 1.7 GB
    of pinned memory per rank) -- the LIST is one million long and every chunk is uploaded, prepared
    and searched anew; only the host buffers repeat."""
    import torch
    import torch.distributed as dist
    from enzymm_b200.engine import Engine, HIT_DTYPE, PackedBatch, Session
    from enzymm_b200.library import CompiledLibrary
    from enzymm_b200.sharding import ChunkQueue, gather_hit_buffer
    from enzymm_b200.synth import kind_classes

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    chunk = args.chunk
    distinct_chunks = max(1, args.distinct // chunk)
    workers = max(1, host_threads() // max(world, 1))
    templates = active_templates()
    dists = [DEFAULT_DIST[min(t.effective_size, 8)] for t in templates]
    engine = Engine(CompiledLibrary(templates, 2.0, dists, dists), device=local_rank)

    def pinned(a):
        return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    # distinct chunk c is generated by rank c mod N and broadcast: nobody generates the corpus twice
    t_gen = time.perf_counter()
    host_chunks = []
    for c in range(distinct_chunks):
        src = c % world
        cols = None
        if rank == src:
            made = make_workload(c, chunk, args.residues, args.chains, workers)
            # atom KINDS travel, every rank classifies them through its own compiled library
            cols = [made.atom_off, made.xyz, made.kind.astype(np.int16), made.residue, made.bfactor]
        if world > 1:
            shapes = [[(a.shape, str(a.dtype)) for a in cols]] if rank == src else [None]
            dist.broadcast_object_list(shapes, src=src)
            received = []
            for i, (shape, dtype) in enumerate(shapes[0]):
                n_bytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
                buf = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
                if rank == src:
                    buf.copy_(torch.from_numpy(np.ascontiguousarray(cols[i]).view(np.uint8).reshape(-1)))
                dist.broadcast(buf, src=src)
                received.append(buf.cpu().numpy().view(dtype).reshape(shape))
            cols = received
        host_chunks.append(PackedBatch(pinned(cols[0]), pinned(cols[1]), pinned(kind_classes(engine.compiled)[cols[2]]),
                                       pinned(cols[3]), pinned(cols[4]), None, None))
    t_gen = time.perf_counter() - t_gen
    max_atoms = max(b.n_atoms for b in host_chunks)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]          # kept alive: the sessions only hold the raw handles
    lanes = [(Session(engine.device_library, max_atoms, chunk, 64 * chunk), st.cuda_stream) for st in streams]
    kwargs = dict(max_candidates=10000, ignore_chain=True, reset=True, force_prepare=True)
    # every chunk's hits are downloaded back to back into ONE pinned buffer (no per-chunk copies; the
    # merge moves it with a single transfer): room for 12 hits per structure of this rank's share
    hit_room = int(12 * (args.total / world) * 1.25) + 64 * chunk
    hit_tensor = torch.empty(hit_room * HIT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    hit_records = hit_tensor.numpy().view(HIT_DTYPE)

    def sweep(n_total):
        """Search chunks of an n_total-long list until the queue is empty; returns the blocks
        [(first structure, offset in hit_records, count)] (structure indices made global)."""
        queue = ChunkQueue(n_total, chunk)
        spans = iter(queue)
        blocks, flying = [], []
        used = 0

        def submit(span, lane):
            sess, st = lanes[lane]
            batch = host_chunks[(span[0] // chunk) % distinct_chunks]
            if span[1] - span[0] < batch.n_structures:
                batch = batch.slice(0, span[1] - span[0])
            sess.upload(batch, stream=st)
            sess.run(stream=st, **kwargs)
            return span, lane

        span = next(spans, None)
        i = 0
        while span is not None or flying:
            if span is not None:
                flying.append(submit(span, i % 2))
                i += 1
                span = next(spans, None)
            if len(flying) == 2 or span is None:
                done, lane = flying.pop(0)
                got = lanes[lane][0].download(stream=lanes[lane][1], out=hit_records[used:])
                got["structure"] += done[0]
                blocks.append((done[0], used, len(got)))
                used += len(got)
        return blocks

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    merged_room = None
    if world > 1 and rank == 0:                                # where the merged list lands; pinned once, up front
        merged_room = torch.empty(int(12 * args.total * 1.25) * HIT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    merge = lambda blocks: gather_hit_buffer(hit_tensor, blocks, HIT_DTYPE.itemsize, device="cuda" if world > 1 else None,
                                             out=merged_room)
    for _ in range(max(args.warmup, 1)):                       # warm-up: a short list through the same code
        merge(sweep(2 * world * chunk))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    times, busy, n_hits = [], [], 0
    windows = []
    for _ in range(args.steps):
        barrier()
        w0 = time.time()
        t0 = time.perf_counter()
        blocks = sweep(args.total)
        t_search = time.perf_counter() - t0
        merged = merge(blocks)
        barrier()
        times.append(time.perf_counter() - t0)
        windows.append((w0, time.time()))
        busy.append(t_search)
        if rank == 0:
            merged = merged.numpy().view(HIT_DTYPE)
            n_hits = len(merged)
            order_ok = bool(np.all(np.diff(merged["structure"].astype(np.int64)) >= 0))
            assert order_ok and int(merged["structure"].max()) < args.total
        del merged, blocks
    clocks = sampler.stop(windows) if sampler else None
    every_busy = [None] * world
    if world > 1:
        dist.all_gather_object(every_busy, [round(b, 4) for b in busy])
    else:
        every_busy = [[round(b, 4) for b in busy]]
    if rank == 0:
        total_s = sum(times)
        line = base_line(args, world)
        line.update({
            "impl": "b200", "scaling": "strong", "value": args.total * args.steps / total_s,
            "ms_per_step": 1000.0 * total_s / args.steps, "dtype": "f32 (+f64 guard band and superposition)",
            "config": {"workload": f"ONE list of {args.total} synthetic {args.residues}-residue structures (structure i = distinct "
                                   f"structure i mod {distinct_chunks * chunk}; planted M-CSA motifs, seed 20230210) x 6780 active "
                                   f"templates, default --jess thresholds, sharded by query over {world} GPU(s) in chunks of "
                                   f"{chunk} from one shared counter, hit lists merged on rank 0 in input order (BASELINE config 3)",
                       "total_structures": args.total, "chunk": chunk, "distinct_structures": distinct_chunks * chunk,
                       "templates": 6780, "parallelism": f"shard-by-structure x{world}, dynamic chunk queue",
                       "timed_region": "pinned host buffers -> H2D -> prepare -> search -> hits D2H per chunk, then the "
                                       "gather of all hit records to rank 0 and their placement in input order; wall clock "
                                       "between barriers",
                       "l2": "every chunk (~220 MB) exceeds the 126 MB L2"},
            "e2e": {"value": args.total * args.steps / total_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(sum(host_chunks[c % distinct_chunks].nbytes() for c in range((args.total + chunk - 1) // chunk))),
                    "d2h_bytes_per_step": int(n_hits * HIT_DTYPE.itemsize)},
            "hits_per_step": int(n_hits), "merged_in_input_order": True,
            "per_rank_search_s": every_busy, "merge_s": [round(t - max(b[i] for b in every_busy), 4) for i, t in enumerate(times)],
            "clocks": clocks, "generation_s": t_gen, "gpu_launches": int(2 * ((args.total + chunk - 1) // chunk) * args.steps),
        })
        print(json.dumps(line), flush=True)
    for sess, _ in lanes:
        sess.close()
    engine.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--structures", type=int, default=10000, help="structures per GPU per step")
    ap.add_argument("--residues", type=int, default=400)
    ap.add_argument("--chains", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=0, help="structures per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", choices=("weak", "strong"), default="weak",
                    help="strong: ONE list of --total structures over N GPUs, host merge timed (BASELINE config 3)")
    ap.add_argument("--total", type=int, default=1000000, help="length of the list in --scaling strong")
    ap.add_argument("--chunk", type=int, default=2048, help="structures per chunk handed out in --scaling strong")
    ap.add_argument("--distinct", type=int, default=16384, help="distinct synthetic structures behind the list")
    ap.add_argument("--api-files", type=int, default=4096,
                    help="PDB files for the Python-API legs (files -> hits / Match objects / TSV); 0 = skip")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.scaling == "strong":
        run_strong(args, rank, local_rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
