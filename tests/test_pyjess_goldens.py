"""Pins from a real PyJess run, when somebody has harvested them (``tools/harvest_pyjess_goldens.py``
on a machine with ``pip install pyjess enzymm``; this container has no network, so the file does not
exist yet and the test skips).  With ``tests/golden/pyjess_hits.json.gz`` present the oracle must
reproduce every hit PyJess reported -- template, matched atoms in template order, RMSD -- which pins
``match_mode 1``, ``<=`` at the thresholds and the other items SURVEY 8c lists as unpinned."""
import gzip
import json

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from enzymm_b200.structures import Molecule
from enzymm_b200.synth import SynthConfig, generate_batch
from helpers import default_distances

HARVEST = GOLDEN / "pyjess_hits.json.gz"


@pytest.mark.skipif(not HARVEST.exists(), reason="no harvested PyJess goldens (tools/harvest_pyjess_goldens.py)")
def test_oracle_reproduces_harvested_pyjess_hits(active_templates):
    with gzip.open(HARVEST, "rt") as handle:
        data = json.load(handle)
    by_id = {t.id: i for i, t in enumerate(active_templates)}
    chunk = generate_batch(0, data["n_synthetic"], SynthConfig(seed=data["seed"]), active_templates)
    molecules = {"1AMY": Molecule.load(GOLDEN / "1AMY.pdb"),
                 "AF-P0DUB6-F1-model_v4": Molecule.load(GOLDEN / "AF-P0DUB6-F1-model_v4.pdb")}
    for i in range(data["n_synthetic"]):
        molecules[f"synth_{i:07d}"] = chunk.to_molecule(i)
    names = [s["id"] for s in data["structures"]]
    mols = [molecules[n] for n in names]
    dist = np.asarray(default_distances(active_templates))
    raw = oracle.query_raw(mols, oracle.OracleTemplates(active_templates), 2.0, dist, dist, threads=8)
    for si, entry in enumerate(data["structures"]):
        want = {by_id[h["template"]]: h for h in entry["hits"]}
        got = set(np.nonzero(raw[si]["found"])[0].tolist())
        assert got == set(want), (entry["id"], sorted(got ^ set(want))[:5])
        serial = mols[si].column("serial")
        for ti, h in want.items():
            m = len(active_templates[ti])
            assert [int(serial[a]) for a in raw[si, ti]["atoms"][:m]] == [a[0] for a in h["atoms"]], (entry["id"], h["template"])
            assert float(raw[si, ti]["rmsd"]) == pytest.approx(h["rmsd"], abs=1e-6)


@pytest.mark.skipif(not HARVEST.exists(), reason="no harvested PyJess goldens (tools/harvest_pyjess_goldens.py)")
def test_mmcif_reader_reads_what_pyjess_reads(mol_1amy):
    """The harvest holds how the real ``pyjess.Molecule.load`` read an mmCIF rendering of 1AMY (two models,
    auth_* items that differ from the label_* items), with and without ``use_author``."""
    from test_cif_ingest import to_cif
    with gzip.open(HARVEST, "rt") as handle:
        data = json.load(handle)
    if "mmcif" not in data:
        pytest.skip(f"the harvest has no mmCIF section ({data.get('mmcif_error', 'older tool')})")
    text = to_cif(mol_1amy, "1AMY", models=(1, 2))
    for key, use_author in (("label", False), ("auth", True)):
        want = data["mmcif"][key]
        got = Molecule.loads(text, use_author=use_author)
        assert len(got) == want["n_atoms"]
        for i, (serial, name, resname, chain, resnum, x, y, z) in enumerate(want["atoms"]):
            atom = got.atom(i)
            assert (atom.serial, atom.name, atom.residue_name, atom.chain_id, atom.residue_number) == \
                   (serial, name, resname, chain, resnum), (key, i)
            assert (atom.x, atom.y, atom.z) == (x, y, z)
