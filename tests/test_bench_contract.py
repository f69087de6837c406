"""bench.py's reference arm runs on CPU and prints one JSON line with the contract's keys."""
import json
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "structures/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"] and line["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    import os
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_parity_checker_sees_what_it_should():
    """bench.py compares the GPU hits with the oracle hits of its CPU-baseline sample; here the checker
    itself is checked: oracle-derived records agree (0 mismatches), and a moved atom, a changed RMSD bit,
    a flipped verdict, a missing and an extra hit are each counted."""
    import numpy as np

    import bench
    import oracle
    from enzymm_b200.engine import HIT_DTYPE, HIT_ORIENTED, HIT_PASS
    templates = bench.active_templates()
    workload = bench.make_workload(0, 6, 400, 1, 1)
    bench.cpu_oracle_rate(workload, 6, 4)
    raw, mols = bench.cpu_oracle_rate.raw, bench.cpu_oracle_rate.molecules
    rows = []
    for s, t in zip(*np.nonzero(raw["found"])):
        r = raw[s, t]
        m = len(templates[t])
        rec = np.zeros((), dtype=HIT_DTYPE)
        rec["structure"], rec["template_index"], rec["n_atoms"], rec["n_complete"] = s, t, m, r["n_complete"]
        rec["rmsd"], rec["atoms"][:m] = r["rmsd"], r["atoms"][:m]
        moved = (mols[s].xyz[r["atoms"][:m]] - r["qbar"]) @ r["rot"].reshape(3, 3).T + r["tbar"]
        rec["orientation"] = oracle.orientation(templates[t], moved)
        dist = bench.DEFAULT_DIST[min(templates[t].effective_size, 8)]
        ok = oracle.predicted_correct(templates[t].effective_size, dist, float(r["rmsd"]), float(rec["orientation"]))
        rec["flags"] = HIT_ORIENTED | (HIT_PASS if ok else 0)
        rows.append(rec)
    hits = np.array(rows, dtype=HIT_DTYPE)
    assert len(hits) >= 6
    extra = hits[:1].copy()
    extra["structure"] = 7                       # beyond the sample: ignored
    report = bench.parity_against_oracle(np.concatenate([hits, extra]), raw, mols, templates)
    assert report["mismatches"] == 0 and report["hits_gpu"] == report["hits_oracle"] == len(hits)
    assert report["structures"] == 6 and report["pairs"] == 6 * len(templates)

    def broken(edit):
        bad = hits.copy()
        edit(bad)
        return bench.parity_against_oracle(bad, raw, mols, templates)["mismatches"]

    def swap(b): b["atoms"][0][0], b["atoms"][0][1] = b["atoms"][0][1], b["atoms"][0][0]
    def rmsd_bit(b): b["rmsd"][1] = np.nextafter(b["rmsd"][1], 9.0)
    def verdict(b): b["flags"][2] ^= HIT_PASS
    def count(b): b["n_complete"][3] += 1
    assert [broken(f) for f in (swap, rmsd_bit, verdict, count)] == [1, 1, 1, 1]
    assert bench.parity_against_oracle(hits[1:], raw, mols, templates)["mismatches"] == 1
    moved = hits.copy()
    moved["template_index"][0] = (moved["template_index"][0] + 1) % len(templates)
    assert bench.parity_against_oracle(moved, raw, mols, templates)["mismatches"] >= 2


def test_roofline_counters_are_tied_to_the_kernel_build(tmp_path, monkeypatch):
    """Counters captured from another build of the search kernel are reported as stale, not printed."""
    import bench
    data, why = bench.committed_counters()
    assert (data is None) == (why is not None)
    monkeypatch.setattr(bench, "kernel_stamp", lambda: "0" * 64)
    data, why = bench.committed_counters()
    assert data is None and "stale" in why
    assert bench.committed_traffic_bytes(10) is None
    assert bench.committed_issue_figure(1e6, 1.0, 1965.0)["stale"] is True
