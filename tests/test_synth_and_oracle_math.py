import numpy as np
import pytest

import oracle
from enzymm_b200.structures import Molecule
from enzymm_b200.synth import CHUNK, SynthConfig, generate_batch, generate_chunk
from helpers import svd_kabsch


def test_oracle_kabsch_matches_svd():
    rng = np.random.default_rng(7)
    for m in (3, 6, 9, 15, 24):
        for trial in range(20):
            t = rng.normal(size=(m, 3)) * 5 + rng.normal(size=3) * 40
            R0 = np.linalg.qr(rng.normal(size=(3, 3)))[0]
            if np.linalg.det(R0) < 0:
                R0[:, 0] *= -1
            q = (t - t.mean(0)) @ R0.T + rng.normal(size=(m, 3)) * (0.0 if trial == 0 else 0.4) + rng.normal(size=3) * 30
            if trial % 5 == 4:
                q[:, 2] *= -1          # mirror image: the proper-rotation constraint matters
            rmsd, R, qbar, tbar = oracle.kabsch(t, q)
            want, Rw = svd_kabsch(t, q)
            assert rmsd == pytest.approx(want, abs=1e-9)
            assert np.linalg.det(R) == pytest.approx(1.0, abs=1e-9)
            np.testing.assert_allclose(R, Rw, atol=1e-7)


def test_synth_deterministic_and_pdb_roundtrip(active_templates):
    cfg = SynthConfig(n_residues=60)
    a = generate_chunk(3, cfg, active_templates, 8)
    b = generate_chunk(3, cfg, active_templates, 5)
    n5 = int(a.atom_off[5])
    assert np.array_equal(a.xyz[:n5], b.xyz) and np.array_equal(a.kind[:n5], b.kind)
    assert [p for p in a.planted if p[0] < 5] == b.planted
    other = generate_chunk(4, cfg, active_templates, 5)
    assert not np.array_equal(other.xyz[:100], a.xyz[:100])
    # the PDB text parses back to bit-identical doubles, names and residues
    mol = a.to_molecule(2)
    back = Molecule.loads(a.to_pdb(2))
    assert np.array_equal(back.xyz, mol.xyz)
    assert back.column("name").tolist() == mol.column("name").tolist()
    assert back.column("residue_number").tolist() == mol.column("residue_number").tolist()
    assert np.allclose(back.column("temperature_factor"), mol.column("temperature_factor"))
    assert (np.diff(a.residue[:int(a.atom_off[1])]) >= 0).all()


def test_synth_batch_concatenation(active_templates):
    cfg = SynthConfig(n_residues=40, max_motifs=1)
    whole = generate_batch(0, CHUNK + 3, cfg, active_templates[:50])
    assert whole.n_structures == CHUNK + 3
    tail = generate_chunk(1, cfg, active_templates[:50], 3)
    assert np.array_equal(whole.xyz[int(whole.atom_off[CHUNK]):], tail.xyz)
    multi = generate_chunk(0, SynthConfig(n_residues=30, n_chains=3), None, 2)
    assert len(np.unique(multi.chain)) == 3 and multi.residue.max() == 89
