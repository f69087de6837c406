"""Small fixed workload for ncu / stats runs: N synthetic structures vs the full library.
usage: python tools/profile_workload.py [n_structures] [steps] [rank whose structures to use]"""
import os
import sys
import time
from pathlib import Path


ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import DEFAULT_DIST, active_templates, make_workload  # noqa: E402
from enzymm_b200.engine import Engine  # noqa: E402
from enzymm_b200.library import CompiledLibrary  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    templates = active_templates()
    dists = [DEFAULT_DIST[min(t.effective_size, 8)] for t in templates]
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    workload = make_workload(first, n, 400, 1, 8)
    engine = Engine(CompiledLibrary(templates, 2.0, dists, dists, plan_order=os.environ.get("EMM_PLAN", "leaders_first")))
    batch = workload.to_packed(engine.compiled)
    sess = engine.session_for(batch.n_atoms, batch.n_structures)
    sess.upload(batch)
    for i in range(steps):
        t0 = time.perf_counter()
        sess.run(force_prepare=True)
        hits, stats = sess.download(with_stats=True)
        dt = time.perf_counter() - t0
        print(f"step {i}: {n / dt:.1f} structures/s, {len(hits)} hits, kernels ms: prepare {sess.kernel_ms('prepare')[-1]:.2f} "
              f"search {sess.kernel_ms('search')[-1]:.2f}")
    if os.environ.get("EMM_STATS") == "1":
        print("stats", stats)
        c = sess.debug_counters()
        pairs = max(stats["pairs"], 1)
        print("per pair: sweeps %.1f evals %.1f" % (stats["sweeps"] / pairs, stats["dist_evals"] / pairs))
        mask = 2 ** 64 - 1
        us, idx = int(c[0]) >> 24, int(c[0]) & 0xFFFFFF
        first_out, last_out = (~int(c[2])) & mask, int(c[3])
        print(f"slowest structure: {us / 1e3:.2f} ms (index {idx}); slowest pair: {(int(c[5]) >> 24) / 1e3:.2f} ms "
              f"(template {int(c[5]) & 0xFFFFFF}); first CTA out {(last_out - first_out) / 1e6:.2f} ms before the last")
        print(f"warps waiting at the end of a structure: {100.0 * int(c[6]) / max(int(c[7]), 1):.1f} % of warp time")
        print("structure time histogram (2.1 ms buckets, last = more):", [int(v) for v in c[8:20]])
        print("pair time histogram (<1 us, 1, 2-3, 4-7, ... >= 1 ms):", [int(v) for v in c[20:32]])
        print("level: pushed-from-level | validated-alive entering | enter calls   (per pair)")
        for k in range(26):
            print(k, "%.2f | %.2f | %.2f" % (c[32 + k] / pairs, c[64 + k] / pairs, c[96 + k] / pairs))


if __name__ == "__main__":
    main()
