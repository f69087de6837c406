"""Pins the CPU oracle against every golden vector the reference's tests hold for the hot path
(reference tests/test_jess_run.py:75-145 TestMatch, 311-377 TestMatcher; golden PDB/TSV files in
tests/golden/ are the reference's tests/test_data/ fixtures, copied verbatim)."""
import math
import re

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from enzymm_b200.templates import Template, iter_bundle, load_templates
from helpers import oracle_matcher_run

T1_PATH = "5_residues/results/csa3d_0285/csa3d_0285.cluster_1_1_1.1uh3_A396-A262-A356-A471-A472.template.pdb"
T2_PATH = "3_residues/results/csa3d_0415/csa3d_0415.cluster_1_1_2.1be0_A124-A175-A125-A289-A260.template.pdb"


def bundle_templates(paths):
    want = set(paths)
    found = {name: Template.loads(text) for name, text in iter_bundle() if name in want}
    assert set(found) == want
    return [found[p] for p in paths]


def golden_atoms(path):
    rows = []
    for line in (GOLDEN / path).read_text().splitlines():
        if line.startswith("ATOM"):
            rows.append((int(line[6:11]), line[12:16].strip(), line[17:20], int(line[22:26]),
                         float(line[30:38]), float(line[38:46]), float(line[46:54])))
    return rows


@pytest.fixture(scope="module")
def hit1(mol_1amy):
    (t1,) = bundle_templates([T1_PATH])
    hits = oracle.query([mol_1amy], oracle.OracleTemplates([t1]), 2, 1.5, 1.5, max_candidates=10000, ignore_chain=False)[0]
    assert len(hits) == 1
    return t1, hits[0]


@pytest.fixture(scope="module")
def hit2(mol_1amy):
    (t2,) = bundle_templates([T2_PATH])
    hits = oracle.query([mol_1amy], oracle.OracleTemplates([t2]), 2, 1, 1, max_candidates=10000, ignore_chain=False)[0]
    assert len(hits) == 1
    return t2, hits[0]


def test_molecule_counts(mol_1amy, mol_af):
    # test_jess_run.py:75-76, 321-329
    assert len(mol_1amy) == 3339
    assert mol_1amy.id == "1AMY"
    from enzymm_b200.matcher import _COUNTED_RESIDUES

    def residue_count(m):
        keep = np.isin(m.column("residue_name"), list(_COUNTED_RESIDUES))
        return len(np.unique(m.column("residue_number")[keep]))

    assert residue_count(mol_1amy) == 403
    assert residue_count(mol_af) == 511
    assert residue_count(mol_af.conserved(80)) == 494


def test_golden_hit_1(mol_1amy, hit1):
    t1, h = hit1
    assert h.rmsd == pytest.approx(0.32093143, abs=5e-8)                       # test_jess_run.py:77
    txyz = h.transform(mol_1amy.xyz[h.atoms])
    assert oracle.orientation(t1, txyz) == pytest.approx(0.15327054322, abs=5e-8)   # :79
    expected = [(0.2290067979141952, -0.3853409610281773, 0.377114677867322),     # :84-100
                (0.4249816660862038, -0.21966898402981627, -0.3540863184957992),
                (0.45459385444007694, -0.34869961601989985, 0.10687378206512577),
                (-0.8733960645698886, 0.2563504028143271, -0.9840695023070225),
                (-0.510183600042339, -0.1958417994791759, 0.18963368325429997)]
    for got, want in zip(oracle.match_vectors(t1, txyz), expected):
        for a, e in zip(got, want):
            assert math.isclose(a, e, rel_tol=1e-9, abs_tol=1e-9)
    res = [(mol_1amy.column("residue_name")[i], mol_1amy.column("chain_id")[i], int(mol_1amy.column("residue_number")[i]))
           for i in h.atoms[::3]]
    assert res == [("GLU", "A", 204), ("ASP", "A", 87), ("ASP", "A", 179), ("HIS", "A", 288), ("ASP", "A", 289)]  # :112-121


def test_golden_hit_1_atoms_both_frames(mol_1amy, hit1):
    _, h = hit1
    query_frame = golden_atoms("1AMY_matches_no_query.pdb")
    template_frame = golden_atoms("1AMY_matches_template.pdb")
    serials = [int(mol_1amy.column("serial")[i]) for i in h.atoms]
    assert serials == [r[0] for r in query_frame]        # same 15 atoms, same order (OD2 before OD1 for ASP 87/179)
    assert [str(mol_1amy.column("name")[i]) for i in h.atoms] == [r[1] for r in query_frame]
    np.testing.assert_allclose(mol_1amy.xyz[h.atoms], [r[4:] for r in query_frame], atol=1e-9)
    np.testing.assert_allclose(h.transform(mol_1amy.xyz[h.atoms]), [r[4:] for r in template_frame], atol=5.01e-4)


def test_golden_hit_2(mol_1amy, hit2):
    t2, h = hit2
    assert h.rmsd == pytest.approx(1.7353479120, abs=5e-8)                      # test_jess_run.py:133
    txyz = h.transform(mol_1amy.xyz[h.atoms])
    assert oracle.orientation(t2, txyz) == pytest.approx(1.6503123465442575, abs=1e-9)  # :135
    res = [(mol_1amy.column("residue_name")[i], int(mol_1amy.column("residue_number")[i])) for i in h.atoms[::3]]
    assert res == [("TRP", 38), ("HIS", 288), ("ASP", 289)]                       # :142-145
    assert [str(mol_1amy.column("name")[i]) for i in h.atoms[3:6]] == ["CG", "CD2", "ND1"]


def test_golden_tsv_row(mol_1amy, hit1):
    """results.tsv: the columns the hot path produces (rmsd, orientation, verdict, residues)."""
    t1, h = hit1
    header, row = [l.split("\t") for l in (GOLDEN / "results.tsv").read_text().splitlines()]
    cell = dict(zip(header, row))
    txyz = h.transform(mol_1amy.xyz[h.atoms])
    orient = oracle.orientation(t1, txyz)
    assert str(round(h.rmsd, 5)) == cell["rmsd"]
    assert str(round(orient, 5)) == cell["orientation"]
    assert str(oracle.predicted_correct(t1.effective_size, 1.5, h.rmsd, orient)) == cell["predicted_correct"]
    assert str(t1.effective_size) == cell["template_effective_size"] and str(t1.dimension) == cell["template_dimension"]


def test_matcher_counts(mol_1amy, mol_af):
    """TestMatcher.test_Matcher_run (test_jess_run.py:301-334): 2, 2, 1, 1, 3."""
    res5 = list(load_templates(subset="5_residues/results/csa3d_0285/"))
    res4 = list(load_templates(subset="4_residues/results/csa3d_0285/"))
    res3 = list(load_templates(subset="3_residues/results/csa3d_0344/"))
    run1 = oracle_matcher_run(res5 + res4, [mol_1amy, mol_af])
    assert list(run1) == [0, 1]
    assert [len(run1[0]), len(run1[1])] == [2, 2]
    run2 = oracle_matcher_run(res5 + res4, [mol_1amy, mol_af.conserved(80)], skip_smaller_hits=True)
    assert [len(run2[0]), len(run2[1])] == [1, 1]
    run3 = oracle_matcher_run(res5 + res4 + res3, [mol_1amy], match_small_templates=True)
    assert len(run3[0]) == 3


SIX = [
    "5_residues/results/csa3d_0285/csa3d_0285.cluster_1_1_1.1uh3_A396-A262-A356-A471-A472.template.pdb",
    "5_residues/results/csa3d_0045/csa3d_0045.cluster_1_1_1.2cxg_A227-A229-A257-A327-A328.template.pdb",
    "3_residues/results/csa3d_0285/csa3d_0285.cluster_1_1_2.1uh3_A396-A262-A356-A471-A472.template.pdb",
    "3_residues/results/csa3d_0285/csa3d_0285.cluster_1_2_2.1uh3_A396-A262-A356-A471-A472.template.pdb",
    "3_residues/results/csa3d_0421/csa3d_0421.cluster_1_1_3.1bf2_A229-A232-A230-A259-A375-A435-A510-A128.template.pdb",
    "3_residues/results/csa3d_0896/csa3d_0896.cluster_2_1_2.2qy1_A135-A179-A137-A230-A175.template.pdb",
]
SIX_PARAMS = {3: (2, 1.2, 1.2), 4: (2, 1.7, 1.7), 5: (2, 2.0, 2.0), 6: (2, 2.0, 2.0), 7: (2, 2.0, 2.0), 8: (2, 2.0, 2.0)}


def test_matcher_single_run(mol_1amy):
    """TestMatcher.test_Matcher_single_run (test_jess_run.py:336-377)."""
    templates = bundle_templates(SIX)
    unfiltered = oracle_matcher_run(templates, [mol_1amy], jess_params=SIX_PARAMS, filter_matches=False)[0]
    filtered = oracle_matcher_run(templates, [mol_1amy], jess_params=SIX_PARAMS, filter_matches=True)[0]
    assert sorted(m.template.pdb_id for m in filtered) == ["1bf2", "1uh3", "1uh3", "1uh3", "2cxg"]
    assert sorted(m.template.pdb_id for m in unfiltered) == ["1bf2", "1uh3", "1uh3", "1uh3", "2cxg", "2qy1"]
    complete = {(m.template.pdb_id, m.template.cluster.member, m.template.effective_size): m.complete for m in unfiltered}
    assert complete[("1uh3", 1, 5)] and complete[("2cxg", 1, 5)]
    assert complete[("1uh3", 1, 3)] and complete[("1uh3", 2, 3)]
    assert not complete[("2qy1", 1, 3)] and not complete[("1bf2", 1, 3)]


def test_full_library_regression(mol_1amy, mol_af, active_templates):
    """Restatement regression targets: 13 / 11 raw hits, 120 / 68 complete assignments,
    11 / 6 pass the filter, no pair at the 10 000-candidate cap."""
    from helpers import default_distances
    dist = np.asarray(default_distances(active_templates))
    raw = oracle.query_raw([mol_1amy, mol_af], oracle.OracleTemplates(active_templates), 2.0, dist, dist, threads=8)
    assert raw["found"].sum(axis=1).tolist() == [13, 11]
    assert raw["n_complete"].sum(axis=1).tolist() == [120, 68]
    assert int(raw["overflow"].sum()) == 0
    passing = []
    for mi, mol in enumerate((mol_1amy, mol_af)):
        n = 0
        for ti in np.nonzero(raw[mi]["found"])[0]:
            r = raw[mi, ti]
            t = active_templates[ti]
            atoms = r["atoms"][:len(t)]
            xyz = (mol.xyz[atoms] - r["qbar"]) @ r["rot"].reshape(3, 3).T + r["tbar"]
            n += oracle.predicted_correct(t.effective_size, dist[ti], float(r["rmsd"]), oracle.orientation(t, xyz))
        passing.append(n)
    assert passing == [11, 6]
