"""Diagnostic: CUDA engine vs CPU oracle on the two fixture structures, full shipped library.

Run on a GPU box:  python tools/gpu_check.py [n_templates]
Prints every disagreement (hit set, atoms, RMSD, orientation, filter verdict, candidate count).
"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from enzymm_b200.engine import Engine, HIT_PASS  # noqa: E402
from enzymm_b200.library import CompiledLibrary  # noqa: E402
from enzymm_b200.packing import pack_molecules  # noqa: E402
from enzymm_b200.structures import Molecule  # noqa: E402
from enzymm_b200.templates import load_templates  # noqa: E402

PARAMS = {3: 0.9, 4: 1.7, 5: 2.0, 6: 2.0, 7: 2.0, 8: 2.0}


def main():
    limit = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    templates = [t for t in load_templates() if t.effective_size >= 3]
    if limit:
        step = max(1, len(templates) // limit)
        templates = templates[::step][:limit]
    dist = [PARAMS[min(t.effective_size, 8)] for t in templates]
    t0 = time.time()
    compiled = CompiledLibrary(templates, 2.0, dist, dist)
    print(f"compiled {len(templates)} templates in {time.time() - t0:.2f}s: ttypes={compiled.n_ttype} "
          f"classes={compiled.n_classes} leaders={len(compiled.leader_ttype)}")
    mols = [Molecule.load(ROOT / "tests/golden/1AMY.pdb"), Molecule.load(ROOT / "tests/golden/AF-P0DUB6-F1-model_v4.pdb")]
    engine = Engine(compiled, 0)
    batch = pack_molecules(mols, compiled)
    t0 = time.time()
    hits, stats = engine.query(batch, max_candidates=10000, ignore_chain=True, with_stats=True)
    t_gpu = time.time() - t0
    t0 = time.time()
    hits2 = engine.query(batch, max_candidates=10000, ignore_chain=True)
    t_gpu2 = time.time() - t0
    print(f"gpu: {len(hits)} hits in {t_gpu:.3f}s (second call {t_gpu2:.3f}s) stats={stats}")
    assert len(hits2) == len(hits)

    ot = oracle.OracleTemplates(templates)
    t0 = time.time()
    raw = oracle.query_raw(mols, ot, 2.0, np.asarray(dist), np.asarray(dist), max_candidates=10000,
                           ignore_chain=True, threads=os.cpu_count() or 1)
    print(f"oracle: {int(raw['found'].sum())} hits in {time.time() - t0:.2f}s; complete={int(raw['n_complete'].sum())}")

    gpu = {(int(h["structure"]), int(h["template_index"])): h for h in hits}
    bad = 0
    for mi in range(len(mols)):
        for ti in range(len(templates)):
            r = raw[mi, ti]
            h = gpu.get((mi, ti))
            if bool(r["found"]) != (h is not None):
                bad += 1
                print(f"MISMATCH found: mol {mi} tpl {ti} ({templates[ti].template_id_string}) oracle={bool(r['found'])} "
                      f"gpu={h is not None} oracle_complete={int(r['n_complete'])}")
                continue
            if h is None:
                continue
            m = len(templates[ti])
            oa = [int(v) for v in r["atoms"][:m]]
            ga = [int(v) for v in h["atoms"][:m]]
            t = templates[ti]
            hit = oracle.OracleHit(mi, ti, float(r["rmsd"]), oa, r["rot"].reshape(3, 3), r["qbar"], r["tbar"],
                                   int(r["n_complete"]), int(r["n_accepted"]), bool(r["overflow"]))
            o_orient = oracle.orientation(t, hit.transform(mols[mi].xyz[oa]))
            o_pass = oracle.predicted_correct(t.effective_size, dist[ti], hit.rmsd, o_orient)
            g_pass = bool(int(h["flags"]) & HIT_PASS)
            problems = []
            if oa != ga:
                problems.append(f"atoms oracle={oa} gpu={ga}")
            if float(h["rmsd"]) != float(r["rmsd"]):
                problems.append(f"rmsd oracle={float(r['rmsd'])!r} gpu={float(h['rmsd'])!r}")
            if abs(float(h["orientation"]) - o_orient) > 1e-9:
                problems.append(f"orientation oracle={o_orient!r} gpu={float(h['orientation'])!r}")
            if g_pass != o_pass:
                problems.append(f"pass oracle={o_pass} gpu={g_pass}")
            if int(h["n_complete"]) != int(r["n_complete"]):
                problems.append(f"n_complete oracle={int(r['n_complete'])} gpu={int(h['n_complete'])}")
            if not np.allclose(h["rot"], r["rot"], atol=1e-12):
                problems.append("rot differs")
            if problems:
                bad += 1
                print(f"MISMATCH mol {mi} tpl {ti} ({t.template_id_string}): " + "; ".join(problems))
    print(f"RESULT: {bad} disagreements over {len(mols) * len(templates)} pairs; gpu hits per mol: "
          f"{[sum(1 for k in gpu if k[0] == mi) for mi in range(len(mols))]}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
