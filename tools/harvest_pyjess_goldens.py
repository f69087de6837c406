"""Harvest golden hit lists from the REAL reference stack -- run this on any machine where
``pip install pyjess enzymm`` works (this build container has neither: no network).

    python tools/harvest_pyjess_goldens.py [n_synthetic] > /dev/null
    -> tests/golden/pyjess_hits.json.gz           (commit it; tests/test_pyjess_goldens.py consumes it)

For the two fixture structures and ``n_synthetic`` (default 256) synthetic 400-residue structures of
the bench generator (seed 20230210, written as PDB text so PyJess reads exactly what this repo's
reader reads) it runs what ``enzymm.jess_run.Matcher._run_jess`` runs (``jess_run.py:800-811``) --
``pyjess.Jess(templates).query(molecule, rmsd, distance, max_dyn, max_candidates=10000,
best_match=True, ignore_chain=True)`` per effective-size group at the default thresholds -- and
records every hit: template id, matched atoms in template order (serial, name, residue name, chain,
residue number), ``rmsd`` and ``log_evalue``.  That pins what no vector in the reference's tests
pins: the reading of ``match_mode 1``, ``<=`` at the thresholds, ``log_evalue``.
It also records how the real ``Molecule.load`` reads an mmCIF rendering of 1AMY (``use_author`` both ways),
which pins this repo's mmCIF reader (``tests/test_pyjess_goldens.py``).
"""
import gzip
import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

DEFAULTS = {3: (2, 0.9, 0.9), 4: (2, 1.7, 1.7), 5: (2, 2.0, 2.0), 6: (2, 2.0, 2.0), 7: (2, 2.0, 2.0), 8: (2, 2.0, 2.0)}


def main():
    import pyjess                                   # the real one
    from enzymm import template as ref_template     # the real EnzyMM

    from enzymm_b200.synth import SynthConfig, generate_batch
    from enzymm_b200.templates import load_templates as own_templates

    n_synth = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    templates = [t for t in ref_template.load_templates(None, with_annotations=False) if t.effective_size >= 3]
    groups = {}
    for t in templates:
        groups.setdefault(min(t.effective_size, 8), []).append(t)
    own_active = [t for t in own_templates() if t.effective_size >= 3]        # plants the same motifs as bench.py
    chunk = generate_batch(0, n_synth, SynthConfig(), own_active)
    work = tempfile.mkdtemp(prefix="emm_harvest_")
    paths = [ROOT / "tests" / "golden" / "1AMY.pdb", ROOT / "tests" / "golden" / "AF-P0DUB6-F1-model_v4.pdb"]
    for i in range(n_synth):
        path = Path(work) / f"synth_{i:07d}.pdb"
        path.write_text(chunk.to_pdb(i))
        paths.append(path)
    out = {"pyjess_version": pyjess.__version__, "n_synthetic": n_synth, "seed": SynthConfig().seed,
           "parameters": {str(k): v for k, v in DEFAULTS.items()}, "structures": []}
    for path in paths:
        molecule = pyjess.Molecule.load(str(path), id=path.stem)
        hits = []
        for size in sorted(groups, reverse=True):
            rmsd, dist, dyn = DEFAULTS[size]
            query = pyjess.Jess(groups[size]).query(molecule, rmsd, dist, dyn, max_candidates=10000, best_match=True,
                                                    ignore_chain=True)
            for hit in query:
                hits.append({"template": hit.template.id, "effective_size": hit.template.effective_size,
                             "rmsd": hit.rmsd, "log_evalue": hit.log_evalue,
                             "atoms": [[a.serial, a.name, a.residue_name, a.chain_id, a.residue_number]
                                       for a in hit.atoms(transform=False)]})
        out["structures"].append({"id": path.stem, "n_atoms": len(molecule), "hits": hits})
        print(path.stem, len(hits), file=sys.stderr)
    # how PyJess reads mmCIF (this repo's choices -- first model, label_* identifiers unless use_author, where
    # '.' falls back to auth_* -- are unpinned: the reference holds no mmCIF input): an mmCIF rendering of
    # 1AMY whose auth_* items differ from its label_* items, through the real Molecule.load both ways
    sys.path.insert(0, str(ROOT / "tests"))
    try:
        from test_cif_ingest import to_cif
        from enzymm_b200.structures import Molecule as OwnMolecule
        own = OwnMolecule.load(ROOT / "tests" / "golden" / "1AMY.pdb")
        cif_path = Path(work) / "1AMY.cif"
        cif_path.write_text(to_cif(own, "1AMY", models=(1, 2)))
        out["mmcif"] = {}
        for use_author in (False, True):
            try:
                mol = pyjess.Molecule.load(str(cif_path), format="detect", use_author=use_author)
            except TypeError:                        # a PyJess without the keyword
                mol = pyjess.Molecule.load(str(cif_path))
            out["mmcif"]["auth" if use_author else "label"] = {
                "id": mol.id, "n_atoms": len(mol),
                "atoms": [[a.serial, a.name, a.residue_name, a.chain_id, a.residue_number, a.x, a.y, a.z]
                          for a in list(mol)[:200]]}
    except Exception as exc:                        # noqa: BLE001 -- e.g. a PyJess built without gemmi
        out["mmcif_error"] = f"{type(exc).__name__}: {exc}"
    target = ROOT / "tests" / "golden" / "pyjess_hits.json.gz"
    with gzip.open(target, "wt") as handle:
        json.dump(out, handle)
    print(f"wrote {target}", file=sys.stderr)


if __name__ == "__main__":
    main()
