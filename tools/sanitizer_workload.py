"""Small workload that reaches every instantiation and rare path of the search kernel, for
``compute-sanitizer --tool racecheck|memcheck`` (logs kept under profiles/).

    compute-sanitizer --tool racecheck python tools/sanitizer_workload.py

Covers: staged blobs <0,1,0>; a 4-chain assembly searched in place <0,0,0>; the cell-list expansion
<0,1,1> / <0,0,1>; skip-mode launches; loose cutoffs with queue overflow, the candidate cap and long
root lists (general level-0 path); dynamic distances and the chain rule (FP64 side paths).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import DEFAULT_DIST, active_templates  # noqa: E402
from enzymm_b200.engine import Engine  # noqa: E402
from enzymm_b200.library import CompiledLibrary  # noqa: E402
from enzymm_b200.synth import SynthConfig, generate_chunk  # noqa: E402


def main():
    step = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    light = len(sys.argv) > 2 and sys.argv[2] == "light"      # racecheck is ~100x slower: bounded blow-ups
    templates = active_templates()[::step]
    dists = [DEFAULT_DIST[min(t.effective_size, 8)] for t in templates]
    small = generate_chunk(0, SynthConfig(), templates, 3)
    big = generate_chunk(0, SynthConfig(n_chains=4), templates, 1)
    two = generate_chunk(1, SynthConfig(n_residues=150, n_chains=2), templates, 2)
    total = 0
    engine = Engine(CompiledLibrary(templates, 2.0, dists, dists))
    for label, chunk, kwargs in (("staged", small, {}), ("in place", big, {}), ("cells staged", small, {"cell_threshold": 1}),
                                 ("cells in place", big, {"cell_threshold": 1}),
                                 ("masked", small, {"conservation_cutoff": 70.0})):
        hits = engine.query(chunk.to_packed(engine.compiled), **kwargs)
        print(label, len(hits), "hits")
        total += len(hits)
    sess = engine.session_for(small.n_atoms, small.n_structures)
    sess.upload(small.to_packed(engine.compiled))
    half = len(templates) // 2
    sess.run(template_begin=0, template_end=half, skip_mode=2, reset=True, force_prepare=True)
    sess.run(template_begin=half, template_end=len(templates), skip_mode=2, reset=False)
    print("skip mode", len(sess.download()), "hits")
    # pair splitting: few templates (idle warps from the start), donation at every level entry
    few = Engine(CompiledLibrary(templates[:20], 2.0, 3.0, 3.0))
    for chunk in (small, big):
        hits = few.query(chunk.to_packed(few.compiled), max_candidates=2000 if light else 10 ** 6, donate_after=1)
        print("split pairs", len(hits), "hits")
    few.close()
    engine.close()
    loose = [t for t in templates if t.effective_size <= 4][:8 if light else 40]
    engine = Engine(CompiledLibrary(loose, 2.0, 3.0, 3.0))
    for cap in ((2000, 50) if light else (10 ** 7, 50)):
        hits = engine.query(small.to_packed(engine.compiled), max_candidates=cap)
        print("loose cutoff, cap", cap, len(hits), "hits, most complete assignments", int(hits["n_complete"].max()))
    engine.close()
    engine = Engine(CompiledLibrary(templates[:60], 2.0, 1.0, 2.5))
    batch = two.to_packed(engine.compiled, with_chain=True)
    for ignore in (True, False):
        hits = engine.query(batch, ignore_chain=ignore)
        print("dynamic distances, ignore_chain", ignore, len(hits), "hits")
    engine.close()
    print("done", total)


if __name__ == "__main__":
    main()
