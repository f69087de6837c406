"""Search time per template size group (one launch per group) on the bench workload.
usage: python tools/size_group_times.py [n_structures] [rank]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import DEFAULT_DIST, active_templates, make_workload  # noqa: E402
from enzymm_b200.engine import Engine  # noqa: E402
from enzymm_b200.library import CompiledLibrary  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    templates = sorted(active_templates(), key=lambda t: t.effective_size)     # stable: contiguous size groups
    sizes = np.asarray([t.effective_size for t in templates])
    atoms = np.asarray([len(t) for t in templates])
    dists = [DEFAULT_DIST[min(t.effective_size, 8)] for t in templates]
    engine = Engine(CompiledLibrary(templates, 2.0, dists, dists))
    batch = make_workload(rank, n, 400, 1, 8).to_packed(engine.compiled)
    sess = engine.session_for(batch.n_atoms, batch.n_structures)
    sess.upload(batch)
    sess.run(force_prepare=True)
    sess.download()
    sess.clear_timings()
    sess.run()
    sess.download()
    total = sess.kernel_ms("search")[-1]
    print(f"all {len(templates)} templates: {total:.1f} ms for {n} structures")
    order_ok = bool(np.all(np.diff(sizes) >= 0) or np.all(np.diff(sizes) <= 0))
    print("library order is monotone in size:", order_ok)
    for size in sorted(set(sizes.tolist())):
        idx = np.nonzero(sizes == size)[0]
        lo, hi = int(idx[0]), int(idx[-1]) + 1
        if hi - lo != len(idx):
            print(f"size {size}: not contiguous, skipped")
            continue
        sess.clear_timings()
        sess.run(template_begin=lo, template_end=hi)
        hits = sess.download()
        ms = sess.kernel_ms("search")[-1]
        print(f"size {size}: templates [{lo},{hi}) n={hi - lo} atoms/template {atoms[idx].mean():.1f} delta {dists[lo]}: "
              f"{ms:.1f} ms ({100 * ms / total:.1f} % of the full run), {len(hits)} hits, {1e3 * ms / (hi - lo) / n:.2f} us/pair")


if __name__ == "__main__":
    main()
