"""File screening throughput: N synthetic PDB files on local disk -> hits.
usage: python tools/ingest_bench.py [n_files] [chunk_size]
Prints (a) native files -> packed columns rate (host only), (b) Molecule path rate on a sample,
(c) Matcher.scan_files end to end (ingest overlapped with the GPU search)."""
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import active_templates, make_workload  # noqa: E402
from enzymm_b200 import jess_run  # noqa: E402
from enzymm_b200.packing import pack_files, pack_molecules  # noqa: E402
from enzymm_b200.structures import load_many  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    templates = active_templates()
    distinct = min(n, int(os.environ.get("EMM_DISTINCT", "1024")))     # file contents repeat beyond this
    workload = make_workload(0, distinct, 400, 1, 8)
    d = tempfile.mkdtemp(prefix="emm_ingest_")
    paths = []
    t0 = time.perf_counter()
    texts = [workload.to_pdb(i).encode() for i in range(distinct)]
    for i in range(n):
        p = os.path.join(d, f"s{i:06d}.pdb")
        with open(p, "wb") as f:
            f.write(texts[i % distinct])
        paths.append(p)
    size = sum(os.path.getsize(p) for p in paths)
    print(f"wrote {n} files, {size / 1e6:.0f} MB in {time.perf_counter() - t0:.1f} s; host threads {len(os.sched_getaffinity(0))}")
    matcher = jess_run.Matcher(templates=templates)
    engine = matcher._ensure_engine()
    for rep in range(2):
        t0 = time.perf_counter()
        batch, _ = pack_files(paths, engine.compiled)
        dt = time.perf_counter() - t0
        print(f"pack_files: {n / dt:.0f} files/s ({size / dt / 1e9:.2f} GB/s of PDB text)")
    sample = paths[:min(n, 256)]
    t0 = time.perf_counter()
    pack_molecules(load_many(sample), engine.compiled)
    print(f"Molecule path (load_many + pack_molecules): {len(sample) / (time.perf_counter() - t0):.0f} files/s")
    # phase times of one chunk, not overlapped
    t0 = time.perf_counter()
    batch, _ = pack_files(paths[:chunk], engine.compiled)
    t1 = time.perf_counter()
    records = matcher._search(batch)
    t2 = time.perf_counter()
    records = matcher._search(batch)
    t3 = time.perf_counter()
    print(f"one chunk of {chunk}: pack {1e3 * (t1 - t0):.0f} ms, search (upload+run+download) {1e3 * (t2 - t1):.0f} ms, "
          f"again {1e3 * (t3 - t2):.0f} ms; {len(records)} hits")
    for rep in range(3):
        t0 = time.perf_counter()
        hits = passing = 0
        for _, _, records in matcher.scan_files(paths, chunk_size=chunk):
            hits += len(records)
            passing += int(((records["flags"] & 4) != 0).sum())
        dt = time.perf_counter() - t0
        print(f"scan_files: {n / dt:.0f} files/s end to end, {hits} hits, {passing} pass the filter")
    # the reference-shaped path: Molecule objects in, {Molecule: [Match]} out
    sample = paths[:min(n, 2048)]
    t0 = time.perf_counter()
    molecules = jess_run.load_molecules(sample)
    t1 = time.perf_counter()
    matches = matcher.run(molecules)
    t2 = time.perf_counter()
    n_matches = sum(len(v) for v in matches.values())
    print(f"Matcher.run on Molecule objects: load_molecules {len(sample) / (t1 - t0):.0f} files/s, run (pack + search + "
          f"{n_matches} Match objects) {len(sample) / (t2 - t1):.0f} molecules/s, together {len(sample) / (t2 - t0):.0f}/s")
    for p in paths:
        os.unlink(p)
    os.rmdir(d)


if __name__ == "__main__":
    main()
