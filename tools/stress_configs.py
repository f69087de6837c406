"""BASELINE configs 4 and 5 in measurable form (numbers quoted in DESIGN.md).

config 4: --unfiltered, -j 2 3.0 3.0, max_candidates 10^7 (combinatorial blow-up of the backtrack)
config 5: 4 x 400-residue assemblies, conservation cutoff 70 as a real mask, --skip-smaller-hits
usage: python tools/stress_configs.py [n4] [n5]
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import DEFAULT_DIST, active_templates, make_workload  # noqa: E402
from enzymm_b200.engine import Engine, HIT_OVERFLOW, HIT_PASS  # noqa: E402
from enzymm_b200.library import CompiledLibrary  # noqa: E402


CELL = int(os.environ.get("EMM_CELL", "0"))     # > 0: leader lists this long go through the cell list


def config4(n):
    templates = active_templates()
    workload = make_workload(0, n, 400, 1, 8)
    engine = Engine(CompiledLibrary(templates, 2.0, 3.0, 3.0))
    batch = workload.to_packed(engine.compiled)
    sess = engine.session_for(batch.n_atoms, batch.n_structures, hit_capacity=8192 * max(n, 1))
    sess.upload(batch)
    for cap in (10000, 10 ** 7):
        t0 = time.perf_counter()
        sess.run(max_candidates=cap, force_prepare=True)
        hits = sess.download()
        dt = time.perf_counter() - t0
        over = int(((hits["flags"] & HIT_OVERFLOW) != 0).sum())
        print(f"config4 max_candidates={cap}: {n} structures in {dt:.2f}s = {n / dt:.2f} structures/s; hits {len(hits)} "
              f"({len(hits) / n:.0f}/structure), overflow-flagged {over}, complete assignments examined "
              f"{int(hits['n_complete'].astype(np.int64).sum()):,} (max/pair {int(hits['n_complete'].max()) if len(hits) else 0})", flush=True)
    engine.close()


def config5(n):
    templates = active_templates()
    dists = [DEFAULT_DIST[min(t.effective_size, 8)] for t in templates]
    order = np.argsort([-t.effective_size for t in templates], kind="stable")
    templates = [templates[i] for i in order]
    dists = [dists[i] for i in order]
    sizes = np.array([t.effective_size for t in templates])
    workload = make_workload(0, n, 400, 4, 8)
    engine = Engine(CompiledLibrary(templates, 2.0, dists, dists))
    batch = workload.to_packed(engine.compiled)
    sess = engine.session_for(batch.n_atoms, batch.n_structures)
    sess.upload(batch)
    for label, cutoff, skip in (("no mask, all sizes", 0.0, False), ("mask 70 + skip-smaller", 70.0, True)):
        t0 = time.perf_counter()
        if skip:
            first = True
            for size in sorted(set(sizes.tolist()), reverse=True):
                idx = np.nonzero(sizes == size)[0]
                sess.run(conservation_cutoff=cutoff, template_begin=int(idx[0]), template_end=int(idx[-1]) + 1,
                         skip_mode=1, reset=first, force_prepare=first, cell_threshold=CELL)
                first = False
        else:
            sess.run(conservation_cutoff=cutoff, force_prepare=True, cell_threshold=CELL)
        hits, stats = sess.download(with_stats=True)
        dt = time.perf_counter() - t0
        print(f"config5 [cell_threshold={CELL}] {label}: {n} assemblies ({batch.n_atoms / n:.0f} atoms each) in {dt:.2f}s = {n / dt:.1f} structures/s; "
              f"hits {len(hits)}, passing {int(((hits['flags'] & HIT_PASS) != 0).sum())}, kept atoms/structure "
              f"{stats['kept_atoms'] / n:.0f}", flush=True)
    engine.close()


if __name__ == "__main__":
    n4 = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n5 = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    if n5:
        config5(n5)
    if n4:
        config4(n4)
