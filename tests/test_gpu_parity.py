"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's goldens.

Bar (BASELINE.json north_star): the match set -- template, query atom indices -- is bit-exact; RMSD
and orientation within 1e-4 (in fact RMSD is bit-identical: both sides evaluate the same FP64
expressions); post-filter decisions identical.
"""
import io
import math

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from enzymm_b200 import jess_run, pyjess
from enzymm_b200.engine import (Engine, EngineError, HIT_NO_MODEL, HIT_OVERFLOW, HIT_PASS, PackedBatch)
from enzymm_b200.library import CompiledLibrary
from enzymm_b200.packing import pack_molecules
from enzymm_b200.structures import Molecule
from enzymm_b200.synth import SynthConfig, generate_chunk
from enzymm_b200.templates import load_templates
from helpers import DEFAULT_PARAMS, default_distances, oracle_matcher_run
from test_oracle_golden import SIX, SIX_PARAMS, T1_PATH, T2_PATH, bundle_templates

pytestmark = pytest.mark.gpu

RMSD_TOL = 1e-4      # stated tolerance of north_star (Angstrom / radians)


def compare_with_oracle(engine, templates, molecules, dist, *, rmsd=2.0, max_candidates=10000,
                        ignore_chain=True, cutoff=0.0, oracle_molecules=None, atom_maps=None, donate_after=0):
    """Run both sides on the same inputs and assert the north_star bar.  Returns the GPU hits."""
    batch = pack_molecules(molecules, engine.compiled)
    hits = engine.query(batch, max_candidates=max_candidates, ignore_chain=ignore_chain,
                        conservation_cutoff=cutoff, donate_after=donate_after)
    dist = np.broadcast_to(np.asarray(dist, dtype=np.float64), (len(templates),)).copy()
    omols = oracle_molecules if oracle_molecules is not None else molecules
    raw = oracle.query_raw(omols, oracle.OracleTemplates(templates), rmsd, dist, dist,
                           max_candidates=max_candidates, ignore_chain=ignore_chain, threads=8)
    gpu = {(int(h["structure"]), int(h["template_index"])): h for h in hits}
    for mi in range(len(molecules)):
        for ti in range(len(templates)):
            r, h = raw[mi, ti], gpu.get((mi, ti))
            if r["overflow"] or (h is not None and int(h["flags"]) & HIT_OVERFLOW):
                # enumeration-order dependent by definition: both sides must agree that the cap was hit
                assert h is None or bool(int(h["flags"]) & HIT_OVERFLOW) == bool(r["overflow"])
                continue
            assert bool(r["found"]) == (h is not None), (mi, ti, templates[ti].template_id_string)
            if h is None:
                continue
            m = len(templates[ti])
            want = [int(v) for v in r["atoms"][:m]]
            if atom_maps is not None:        # oracle ran on a masked copy: map back to original indices
                want = [int(atom_maps[mi][v]) for v in want]
            assert [int(v) for v in h["atoms"][:m]] == want, (mi, ti)
            assert float(h["rmsd"]) == float(r["rmsd"])                 # bit-identical, well inside RMSD_TOL
            assert abs(float(h["rmsd"]) - float(r["rmsd"])) <= RMSD_TOL
            assert int(h["n_complete"]) == int(r["n_complete"])
            t = templates[ti]
            if getattr(t, "residues", None):
                xyz = (omols[mi].xyz[[int(v) for v in r["atoms"][:m]]] - r["qbar"]) @ r["rot"].reshape(3, 3).T + r["tbar"]
                o_orient = oracle.orientation(t, xyz)
                assert abs(float(h["orientation"]) - o_orient) <= RMSD_TOL
                try:
                    o_pass = oracle.predicted_correct(t.effective_size, dist[ti], float(r["rmsd"]), o_orient)
                except KeyError:
                    assert int(h["flags"]) & HIT_NO_MODEL
                else:
                    assert bool(int(h["flags"]) & HIT_PASS) == o_pass
    return hits


@pytest.fixture(scope="module")
def full_engine(active_templates):
    dist = default_distances(active_templates)
    eng = Engine(CompiledLibrary(active_templates, 2.0, dist, dist))
    yield eng
    eng.close()


# ---- config 1: fixture structures vs the full shipped library -------------------------------------

def test_fixtures_full_library(full_engine, active_templates, mol_1amy, mol_af):
    hits = compare_with_oracle(full_engine, active_templates, [mol_1amy, mol_af], default_distances(active_templates))
    per_mol = np.bincount(hits["structure"], minlength=2).tolist()
    assert per_mol == [13, 11]
    passing = np.bincount(hits["structure"][(hits["flags"] & HIT_PASS) != 0], minlength=2).tolist()
    assert passing == [11, 6]


def test_reference_golden_through_jess_api(mol_1amy):
    """tests/test_jess_run.py:32-145 of the reference, run against the drop-in API."""
    t1, t2 = bundle_templates([T1_PATH, T2_PATH])
    best = list(pyjess.Jess([t1]).query(mol_1amy, 2, 1.5, 1.5, max_candidates=10000, best_match=True))
    match1 = jess_run.Match(hit=best[0], pairwise_distance=1.5, complete=True, index=0)
    assert match1.hit.molecule().id == "1AMY"
    assert match1.hit.template is t1
    assert match1.query_atom_count == 3339 and match1.query_residue_count == 403
    assert match1.hit.rmsd == pytest.approx(0.32093143, abs=5e-8)
    assert match1.orientation == pytest.approx(0.15327054322, abs=5e-8)
    assert match1.hit.orientation == pytest.approx(match1.orientation, abs=1e-9)      # fused filter value
    expected = [(0.2290067979141952, -0.3853409610281773, 0.377114677867322),
                (0.4249816660862038, -0.21966898402981627, -0.3540863184957992),
                (0.45459385444007694, -0.34869961601989985, 0.10687378206512577),
                (-0.8733960645698886, 0.2563504028143271, -0.9840695023070225),
                (-0.510183600042339, -0.1958417994791759, 0.18963368325429997)]
    for a, e in zip(match1.match_vector_list, expected):
        assert math.isclose(a.x, e[0], rel_tol=1e-9, abs_tol=1e-9)
        assert math.isclose(a.y, e[1], rel_tol=1e-9, abs_tol=1e-9)
        assert math.isclose(a.z, e[2], rel_tol=1e-9, abs_tol=1e-9)
    assert match1.template_vector_list == [r.orientation_vector for r in t1.residues]
    assert match1.preserved_resid_order is True and match1.multimeric is False
    assert match1.matched_residues == [("GLU", "A", "204"), ("ASP", "A", "87"), ("ASP", "A", "179"),
                                       ("HIS", "A", "288"), ("ASP", "A", "289")]
    assert match1.predicted_correct is True and match1.hit.device_pass

    best2 = list(pyjess.Jess([t2]).query(mol_1amy, 2, 1, 1, max_candidates=10000, best_match=True))
    match2 = jess_run.Match(hit=best2[0])
    assert match2.hit.rmsd == pytest.approx(1.7353479120, abs=5e-8)
    assert match2.orientation == pytest.approx(1.6503123465442575, abs=1e-9)
    assert match2.preserved_resid_order is False
    assert match2.matched_residues == [("TRP", "A", "38"), ("HIS", "A", "288"), ("ASP", "A", "289")]


def test_reference_golden_files(mol_1amy):
    """Byte parity of the writers against the reference's golden files (test_jess_run.py:147-179).
    log_evalue and the M-CSA annotation columns are outside the hot path."""
    (t1,) = bundle_templates([T1_PATH])
    hit = next(pyjess.Jess([t1]).query(mol_1amy, 2, 1.5, 1.5, max_candidates=10000, best_match=True))
    match = jess_run.Match(hit=hit, pairwise_distance=1.5, complete=True, index=0)
    for name, kwargs in (("1AMY_matches_no_query.pdb", dict(transform=False, include_query=False)),
                         ("1AMY_matches_query_included.pdb", dict(transform=False, include_query=True)),
                         ("1AMY_matches_template.pdb", dict(transform=True, include_query=False))):
        buffer = io.StringIO()
        match.dump2pdb(buffer, **kwargs)
        assert buffer.getvalue() == (GOLDEN / name).read_text(), name
    got_header, got_row = [l.split("\t") for l in match.dumps(header=True).splitlines()]
    want_header, want_row = [l.split("\t") for l in (GOLDEN / "results.tsv").read_text().splitlines()]
    assert got_header == want_header
    skip = {"log_evalue", "number_of_mutated_residues", "number_of_side_chain_residues_(template,reference)",
            "number_of_metal_ligands_(template,reference)", "number_of_ptm_residues_(template, reference)",
            "total_reference_residues"}
    for col, got, want in zip(want_header, got_row, want_row):
        if col not in skip:
            assert got == want, col


# ---- Matcher semantics (reference TestMatcher) -------------------------------------------------------

def test_matcher_run_counts(mol_1amy, mol_af):
    res5 = list(load_templates(subset="5_residues/results/csa3d_0285/"))
    res4 = list(load_templates(subset="4_residues/results/csa3d_0285/"))
    res3 = list(load_templates(subset="3_residues/results/csa3d_0344/"))
    m1 = jess_run.Matcher(templates=res5 + res4, cpus=2)
    assert m1.template_effective_sizes == [5, 4] and m1.cpus == 2
    mol3 = mol_af.conserved(80)
    out1 = m1.run(molecules=[mol_1amy, mol_af])
    assert list(out1.keys()) == [mol_1amy, mol_af]
    assert len(out1[mol_1amy]) == 2 and len(out1[mol_af]) == 2
    assert [m.query_residue_count for m in out1[mol_af]] == [511, 511]
    out2 = jess_run.Matcher(templates=res5 + res4, skip_smaller_hits=True).run(molecules=[mol_1amy, mol3])
    assert len(out2[mol_1amy]) == 1 and len(out2[mol3]) == 1
    assert [m.query_residue_count for m in out2[mol3]] == [494]
    with pytest.warns(Warning):
        m3 = jess_run.Matcher(templates=res5 + res4 + res3, match_small_templates=True, warn=True, cpus=-1)
    assert len(m3.run(molecules=[mol_1amy])[mol_1amy]) == 3
    with pytest.raises(ValueError):
        jess_run.Matcher(templates=res5 + res5)


def test_matcher_single_run_and_filter(mol_1amy):
    templates = bundle_templates(SIX)
    params = {k: {"rmsd": v[0], "distance": v[1], "max_dynamic_distance": v[2]} for k, v in SIX_PARAMS.items()}
    unfiltered = jess_run.Matcher(templates=templates, jess_params=params, filter_matches=False).run_single(mol_1amy)
    filtered = jess_run.Matcher(templates=templates, jess_params=params, filter_matches=True).run_single(mol_1amy)
    assert sorted(m.hit.template.pdb_id for m in filtered) == ["1bf2", "1uh3", "1uh3", "1uh3", "2cxg"]
    assert sorted(m.hit.template.pdb_id for m in unfiltered) == ["1bf2", "1uh3", "1uh3", "1uh3", "2cxg", "2qy1"]
    want = oracle_matcher_run(templates, [mol_1amy], jess_params=SIX_PARAMS, filter_matches=False)[0]
    assert [(m.hit.template.pdb_id, m.complete) for m in unfiltered] == [(m.template.pdb_id, m.complete) for m in want]
    for m in unfiltered:       # the fused GPU verdict equals the reference formula evaluated in Python
        assert m.hit.device_pass == m.predicted_correct
    # a distance without logistic models is a KeyError, as in the reference (test_cli.py:131-132)
    odd = {k: {"rmsd": 2, "distance": 0.5, "max_dynamic_distance": 0.5} for k in range(3, 9)}
    (t2,) = bundle_templates([T2_PATH])
    loose = {k: {"rmsd": 2, "distance": 1.05, "max_dynamic_distance": 1.05} for k in range(3, 9)}
    with pytest.raises(KeyError):
        jess_run.Matcher(templates=[t2], jess_params=loose).run_single(mol_1amy)
    assert jess_run.Matcher(templates=[t2], jess_params=odd).run_single(mol_1amy) == []


def test_matcher_full_library_vs_oracle(active_templates, mol_1amy, mol_af):
    for kwargs in (dict(), dict(filter_matches=False), dict(skip_smaller_hits=True)):
        got = jess_run.Matcher(templates=active_templates, **kwargs).run([mol_1amy, mol_af])
        want = oracle_matcher_run(active_templates, [mol_1amy, mol_af], threads=8, **kwargs)
        mols = [mol_1amy, mol_af]
        assert [mols.index(k) for k in got.keys()] == list(want.keys())
        for mi, wm in want.items():
            gm = got[mols[mi]]
            assert [m.hit.template.id or m.hit.template.template_id_string for m in gm] == \
                   [m.template.id or m.template.template_id_string for m in wm]
            assert [m.hit.atom_indices for m in gm] == [m.hit.atoms for m in wm]
            assert [m.complete for m in gm] == [m.complete for m in wm]


# ---- config 2 (reduced) and edge cases -----------------------------------------------------------------

def test_synthetic_structures(full_engine, active_templates):
    chunk = generate_chunk(0, SynthConfig(), active_templates, 24)
    mols = [chunk.to_molecule(i) for i in range(chunk.n_structures)]
    hits = compare_with_oracle(full_engine, active_templates, mols, default_distances(active_templates))
    found = {(int(h["structure"]), int(h["template_index"])) for h in hits}
    recovered = sum(1 for p in chunk.planted if p in found)
    assert recovered >= len(chunk.planted) // 2
    # the generator's packed columns are the same input as the Molecule route
    again = full_engine.query(chunk.to_packed(full_engine.compiled))
    assert np.array_equal(again["atoms"], hits["atoms"]) and np.array_equal(again["rmsd"], hits["rmsd"])


def test_empty_and_degenerate_inputs(full_engine, active_templates, mol_1amy):
    gly = Molecule.loads("".join(
        f"ATOM  {i + 1:>5}  CA  GLY A{i + 1:>4}    {i * 3.8:8.3f}{0.0:8.3f}{0.0:8.3f}  1.00 50.00           C\n" for i in range(30)))
    water = Molecule.loads("HETATM    1  O   HOH A 600      -3.288  67.042  32.622  1.00  2.00           O\n")
    mols = [Molecule(), gly, water, mol_1amy, Molecule()]
    hits = compare_with_oracle(full_engine, active_templates, mols, default_distances(active_templates))
    assert set(hits["structure"].tolist()) == {3}
    empty = full_engine.query(pack_molecules([], full_engine.compiled))
    assert len(empty) == 0


def test_conservation_mask_device_equals_host(full_engine, active_templates, mol_af):
    """--conservation-cutoff as a real mask: device-side masking == querying Molecule.conserved()."""
    cutoff = 70.0
    masked = mol_af.conserved(cutoff)
    keep = np.nonzero(mol_af.column("temperature_factor") >= cutoff)[0]
    dist = default_distances(active_templates)
    compare_with_oracle(full_engine, active_templates, [mol_af], dist, cutoff=cutoff,
                        oracle_molecules=[masked], atom_maps=[keep])
    host_side = full_engine.query(pack_molecules([masked], full_engine.compiled))
    dev_side = full_engine.query(pack_molecules([mol_af], full_engine.compiled), conservation_cutoff=cutoff)
    assert np.array_equal(dev_side["template_index"], host_side["template_index"])
    assert np.array_equal(dev_side["rmsd"], host_side["rmsd"])
    assert [keep[a] for a in host_side["atoms"][0][:host_side["n_atoms"][0]]] == \
        dev_side["atoms"][0][:dev_side["n_atoms"][0]].tolist()


def test_translation_invariance_guard_band(active_templates, mol_1amy):
    """Far-from-origin coordinates stress the FP32 guard band: results must not move."""
    subset = active_templates[::9]
    dist = default_distances(subset)
    eng = Engine(CompiledLibrary(subset, 2.0, dist, dist))
    try:
        base = compare_with_oracle(eng, subset, [mol_1amy], dist)
        shifted = mol_1amy.with_xyz(np.round(mol_1amy.xyz + np.array([9000.0, -8000.0, 7000.5]), 3))
        moved = compare_with_oracle(eng, subset, [shifted], dist)
        assert np.array_equal(base["template_index"], moved["template_index"])
        assert np.array_equal(base["atoms"], moved["atoms"])
        np.testing.assert_allclose(base["rmsd"], moved["rmsd"], atol=1e-6)
    finally:
        eng.close()


def test_loose_cutoffs_and_candidate_cap(active_templates, mol_1amy):
    """Config 4 in miniature: loosened distance cutoff; and the max_candidates cap raises OVERFLOW."""
    subset = [t for t in active_templates if t.effective_size == 3][::60]
    eng = Engine(CompiledLibrary(subset, 2.0, 3.0, 3.0))
    try:
        hits = compare_with_oracle(eng, subset, [mol_1amy], 3.0, max_candidates=10 ** 7)
        assert len(hits) > 0 and int(hits["n_complete"].max()) > 50
        capped = compare_with_oracle(eng, subset, [mol_1amy], 3.0, max_candidates=5)
        assert (capped["flags"] & HIT_OVERFLOW).any()
    finally:
        eng.close()


def test_dynamic_distance_and_chain_rule(active_templates):
    """max_dynamic_distance != distance_cutoff (per-pair deltas from distance weights) and
    ignore_chain=False on a two-chain structure -- both unpinned upstream, but GPU == oracle."""
    subset = [t for t in active_templates if t.multimeric][:40] + active_templates[:40]
    chunk = generate_chunk(1, SynthConfig(n_residues=150, n_chains=2), subset, 6)
    mols = [chunk.to_molecule(i) for i in range(chunk.n_structures)]
    for cut, dyn, ignore in ((1.5, 1.5, False), (1.0, 2.5, True), (1.0, 2.5, False)):
        eng = Engine(CompiledLibrary(subset, 2.0, cut, dyn))
        try:
            batch = pack_molecules(mols, eng.compiled)
            hits = eng.query(batch, max_candidates=10000, ignore_chain=ignore)
            dist = np.full(len(subset), cut)
            raw = oracle.query_raw(mols, oracle.OracleTemplates(subset), 2.0, dist, np.full(len(subset), dyn),
                                   max_candidates=10000, ignore_chain=ignore, threads=8)
            got = {(int(h["structure"]), int(h["template_index"])): h for h in hits}
            assert {k for k in got} == {(int(a), int(b)) for a, b in zip(*np.nonzero(raw["found"]))}
            for (mi, ti), h in got.items():
                assert h["atoms"][:len(subset[ti])].tolist() == raw[mi, ti]["atoms"][:len(subset[ti])].tolist()
                assert float(h["rmsd"]) == float(raw[mi, ti]["rmsd"])
        finally:
            eng.close()


def test_large_structure_global_memory_path(full_engine, active_templates):
    """A 4 x 400-residue assembly (config 5 shape) does not fit the shared-memory staging area and
    takes the global-memory path; with pLDDT masking + skip-smaller semantics checked at Matcher level."""
    chunk = generate_chunk(0, SynthConfig(n_chains=4), active_templates, 2)
    mols = [chunk.to_molecule(i) for i in range(2)]
    subset = active_templates[::15]
    dist = default_distances(subset)
    eng = Engine(CompiledLibrary(subset, 2.0, dist, dist))
    try:
        hits, stats = None, None
        batch = pack_molecules(mols, eng.compiled)
        compare_with_oracle(eng, subset, mols, dist)
        masked = [m.conserved(70) for m in mols]
        keeps = [np.nonzero(m.column("temperature_factor") >= 70)[0] for m in mols]
        compare_with_oracle(eng, subset, mols, dist, cutoff=70.0, oracle_molecules=masked, atom_maps=keeps)
    finally:
        eng.close()
    got = jess_run.Matcher(templates=subset, skip_smaller_hits=True, conservation_cutoff=70,
                           apply_conservation_mask=True).run(mols)
    want = oracle_matcher_run(subset, masked, skip_smaller_hits=True, threads=8)
    assert [mols.index(k) for k in got] == list(want)
    for mi, wm in want.items():
        assert [[int(keeps[mi][a]) for a in m.hit.atoms] for m in wm] == [m.hit.atom_indices for m in got[mols[mi]]]


def test_split_residue_is_reordered(active_templates, mol_1amy):
    """Atoms of one residue split over two runs of the file: the host reorders, hits report file indices."""
    order = np.arange(len(mol_1amy))
    asp87 = np.nonzero((mol_1amy.column("residue_number") == 87) & (mol_1amy.column("residue_name") == "ASP"))[0]
    moved = np.concatenate([np.delete(order, asp87[-2:]), asp87[-2:]])      # OD1/OD2 of ASP 87 go to the end
    shuffled = mol_1amy.select(moved)
    (t1,) = bundle_templates([T1_PATH])
    a = next(pyjess.Jess([t1]).query(mol_1amy, 2, 1.5, 1.5, max_candidates=10000, best_match=True))
    b = next(pyjess.Jess([t1]).query(shuffled, 2, 1.5, 1.5, max_candidates=10000, best_match=True))
    assert b.rmsd == pytest.approx(a.rmsd, abs=1e-12)
    assert [int(moved[i]) for i in b.atom_indices] == a.atom_indices


def test_hit_buffer_overflow_is_loud_then_recovers(full_engine, active_templates, mol_1amy):
    from enzymm_b200.engine import Session
    batch = pack_molecules([mol_1amy], full_engine.compiled)
    sess = Session(full_engine.device_library, batch.n_atoms, 1, hit_capacity=4)
    try:
        sess.upload(batch)
        sess.run()
        with pytest.raises(EngineError) as info:
            sess.download()
        assert info.value.status == -4
    finally:
        sess.close()
    assert len(full_engine.query(batch)) == 13


def test_input_contract_violation_is_loud(full_engine):
    """Loud: EMM_ERR_INPUT, with the offender named (see also test_refused_structure_does_not_cost_the_batch)."""
    xyz = np.zeros((4, 3))
    bad = PackedBatch(np.array([0, 4]), xyz, np.ones(4, dtype=np.uint16), np.array([0, 1, 0, 1], dtype=np.int32))
    with pytest.raises(EngineError) as info:
        full_engine.query(bad)
    assert info.value.status == -5


def test_determinism_and_shard_union(full_engine, active_templates):
    """Size-independent properties at a larger size: two runs agree exactly, and searching two
    halves separately gives the same hits as searching the whole batch (what multi-GPU sharding does)."""
    from enzymm_b200.sharding import merge_hits, shard_bounds
    chunk = generate_chunk(2, SynthConfig(), active_templates, 96)
    batch = chunk.to_packed(full_engine.compiled)
    a = full_engine.query(batch)
    b = full_engine.query(batch)
    assert a.tobytes() == b.tobytes()
    parts = []
    for rank in range(2):
        lo, hi = shard_bounds(batch.n_structures, 2, rank)
        parts.append((lo, full_engine.query(batch.slice(lo, hi))))
    merged = merge_hits(parts)
    assert merged.tobytes() == a.tobytes()
    found = {(int(h["structure"]), int(h["template_index"])) for h in a}
    assert sum(1 for p in chunk.planted if p in found) >= 0.5 * len(chunk.planted)


def test_full_size_properties(full_engine, active_templates):
    """Size-independent properties on a batch too large for the oracle (2 560 structures x 6 780
    templates = 17 M pairs): determinism, shard-union, planted-motif recovery, and an independent
    numpy re-derivation of every reported hit (SVD Kabsch RMSD, same-residue rule, thresholds)."""
    from enzymm_b200.synth import generate_batch
    from enzymm_b200.sharding import merge_hits, shard_bounds
    from helpers import svd_kabsch
    work = generate_batch(1024, 2560, SynthConfig(), active_templates)
    batch = work.to_packed(full_engine.compiled)
    hits = full_engine.query(batch)
    assert full_engine.query(batch).tobytes() == hits.tobytes()
    parts = [(lo, full_engine.query(batch.slice(lo, hi)))
             for lo, hi in (shard_bounds(batch.n_structures, 4, r, 256) for r in range(4))]
    assert merge_hits(parts).tobytes() == hits.tobytes()
    found = set(zip(hits["structure"].tolist(), hits["template_index"].tolist()))
    assert len(found) == len(hits)                                   # at most one hit per (structure, template)
    assert sum(1 for p in work.planted if p in found) >= 0.7 * len(work.planted)
    dist = default_distances(active_templates)
    rng = np.random.default_rng(0)
    for h in hits[rng.choice(len(hits), size=400, replace=False)]:
        t = active_templates[int(h["template_index"])]
        m = int(h["n_atoms"])
        lo = int(work.atom_off[int(h["structure"])])
        atoms = h["atoms"][:m] + lo
        assert len(set(atoms.tolist())) == m                         # injective
        q = work.xyz[atoms]
        txyz = np.array([(a.x, a.y, a.z) for a in t])
        rmsd, _ = svd_kabsch(txyz, q)
        assert abs(rmsd - float(h["rmsd"])) < 1e-6 and float(h["rmsd"]) <= 2.0
        res = work.residue[atoms]
        tres = [(a.chain_id, a.residue_number) for a in t]
        for i in range(m):                                           # same template residue -> same query residue
            for j in range(i):
                if tres[i] == tres[j]:
                    assert res[i] == res[j]
        dq = np.linalg.norm(q[:, None] - q[None], axis=-1)
        dt = np.linalg.norm(txyz[:, None] - txyz[None], axis=-1)
        assert np.abs(dq - dt).max() <= dist[int(h["template_index"])] + 1e-9


def test_cell_list_path_equals_list_path(full_engine, active_templates, mol_1amy, mol_af):
    """Leader candidates through the uniform-grid cell list (forced for every leader level with
    cell_threshold=1) give exactly the hits of the typed-list scan and of the oracle -- including
    queue-overflow resumes inside a cell row (loose cutoffs)."""
    chunk = generate_chunk(7, SynthConfig(), active_templates, 6)
    mols = [mol_1amy, mol_af] + [chunk.to_molecule(i) for i in range(chunk.n_structures)]
    batch = pack_molecules(mols, full_engine.compiled)
    by_list = full_engine.query(batch, cell_threshold=-1)
    by_cell = full_engine.query(batch, cell_threshold=1)
    assert by_cell.tobytes() == by_list.tobytes()
    raw = oracle.query_raw(mols, oracle.OracleTemplates(active_templates), 2.0,
                           np.asarray(default_distances(active_templates)), np.asarray(default_distances(active_templates)),
                           threads=8)
    assert {(int(h["structure"]), int(h["template_index"])) for h in by_cell} == \
        {(int(a), int(b)) for a, b in zip(*np.nonzero(raw["found"]))}
    subset = [t for t in active_templates if t.effective_size == 3][::60]
    eng = Engine(CompiledLibrary(subset, 2.0, 3.0, 3.0))
    try:
        loose = pack_molecules([mol_1amy], eng.compiled)
        a = eng.query(loose, max_candidates=10 ** 7, cell_threshold=-1)
        b = eng.query(loose, max_candidates=10 ** 7, cell_threshold=1)
        assert a.tobytes() == b.tobytes() and int(a["n_complete"].max()) > 50
    finally:
        eng.close()
    # a 4-chain assembly searched in place from global memory, cells forced
    big = generate_chunk(0, SynthConfig(n_chains=4), active_templates, 1)
    bb = big.to_packed(full_engine.compiled)
    assert full_engine.query(bb, cell_threshold=1).tobytes() == full_engine.query(bb, cell_threshold=-1).tobytes()


def test_query_batch_entry_point(active_templates, mol_1amy):
    """emm_query_batch -- the one-call C-ABI entry point (upload + run + download with host buffers)
    -- called through ctypes exactly as INTEGRATION.md shows, equals the session path."""
    import ctypes
    from enzymm_b200.engine import DeviceLibrary, HIT_DTYPE, _QueryParams, _Stats, load_cdll
    subset = active_templates[::7]
    dist = default_distances(subset)
    compiled = CompiledLibrary(subset, 2.0, dist, dist)
    dev = DeviceLibrary(compiled)
    try:
        batch = pack_molecules([mol_1amy, mol_1amy.conserved(15)], compiled)
        dev.sync_compat()
        lib = load_cdll()
        hits = np.zeros(256, dtype=HIT_DTYPE)
        n = ctypes.c_int64(0)
        stats = _Stats()
        params = _QueryParams(10000, 1, 0.0, 0, 0, 0, 1, 0, 0, 0)
        st = batch.as_struct()
        rc = lib.emm_query_batch(dev.handle, ctypes.byref(st), ctypes.byref(params),
                                 hits.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(hits)), ctypes.byref(n),
                                 ctypes.byref(stats))
        assert rc == 0, lib.emm_last_error()
        eng = Engine(compiled)
        try:
            want = eng.query(batch)
        finally:
            eng.close()
        assert n.value == len(want) > 0 and hits[:n.value].tobytes() == want.tobytes()
        tiny = np.zeros(1, dtype=HIT_DTYPE)
        rc = lib.emm_query_batch(dev.handle, ctypes.byref(st), ctypes.byref(params),
                                 tiny.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(1), ctypes.byref(n), None)
        assert rc == -4 and n.value == len(want)          # EMM_ERR_CAPACITY reports the required count
    finally:
        dev.close()


def _perturbed_molecule(rng, mol):
    """A structure with the irregularities real files have: deleted atoms, waters / ions / a ligand
    as HETATM, duplicated coordinates, renumbered residues, a second chain, insertion codes."""
    keep = rng.random(len(mol)) > 0.03
    m = mol.select(keep)
    cols = {k: m.column(k).copy() for k in ("serial", "name", "altloc", "residue_name", "chain_id", "residue_number",
                                           "insertion_code", "occupancy", "temperature_factor", "segment",
                                           "element", "charge")}
    xyz = m.xyz.copy()
    n = len(m)
    half = cols["residue_number"] > np.median(cols["residue_number"])
    if rng.random() < 0.5:
        cols["chain_id"][half] = "B"
    if rng.random() < 0.5:
        cols["residue_number"][half] -= int(np.median(cols["residue_number"]))      # numbers repeat across chains
    for _ in range(3):                                                             # duplicated atoms
        i = int(rng.integers(n))
        xyz[int(rng.integers(n))] = xyz[i]
    extra = int(rng.integers(5, 40))                                               # hetero atoms with odd names
    het_names = ["O", "ZN", "MG", "C1'", "O5'", "N1", "CA", "OXT", "FE", "O1A"]
    het_res = ["HOH", "ZN", "MG", "NAD", "HEM", "SO4", "CA", "MSE"]
    ex = {k: [] for k in cols}
    exyz = []
    for j in range(extra):
        nm = het_names[int(rng.integers(len(het_names)))]
        ex["serial"].append(90000 + j); ex["name"].append(nm); ex["altloc"].append(" ")
        ex["residue_name"].append(het_res[int(rng.integers(len(het_res)))]); ex["chain_id"].append("A")
        ex["residue_number"].append(900 + j // 3); ex["insertion_code"].append(" "); ex["occupancy"].append(1.0)
        ex["temperature_factor"].append(20.0); ex["segment"].append(""); ex["element"].append(nm[:1]); ex["charge"].append(0)
        exyz.append(xyz[int(rng.integers(n))] + rng.normal(size=3) * 3.0)
    cols = {k: np.concatenate([cols[k], np.asarray(ex[k], dtype=cols[k].dtype)]) for k in cols}
    xyz = np.round(np.concatenate([xyz, np.asarray(exyz)]), 3)
    return Molecule._from_columns(cols, xyz, mol.id)


def test_randomized_inputs_and_parameters(active_templates):
    """Fuzz: irregular structures x random template subsets x random thresholds, candidate caps and
    chain rule -- the GPU must agree with the oracle everywhere."""
    rng = np.random.default_rng(20230210)
    chunk = generate_chunk(9, SynthConfig(n_residues=220, max_motifs=3), active_templates, 10)
    for trial in range(5):
        mols = [_perturbed_molecule(rng, chunk.to_molecule(int(i))) for i in rng.choice(10, size=3, replace=False)]
        pick = rng.choice(len(active_templates), size=260, replace=False)
        subset = [active_templates[int(i)] for i in sorted(pick)]
        cut = float(rng.choice([0.7, 0.9, 1.3, 1.7, 2.0, 2.6]))
        thr = float(rng.choice([0.6, 1.0, 2.0, 4.0]))
        cap = int(rng.choice([40, 1000, 10000]))
        ignore = bool(rng.integers(2))
        eng = Engine(CompiledLibrary(subset, thr, cut, cut))
        try:
            compare_with_oracle(eng, subset, mols, cut, rmsd=thr, max_candidates=cap, ignore_chain=ignore)
        finally:
            eng.close()


def test_plain_templates_with_odd_residue_groups(mol_1amy):
    """pyjess-level templates (no EnzyMM residues): groups of 1, 2 and 4 atoms, a single-atom
    template, mixed match modes -- exercised through Jess.query against the oracle."""
    from enzymm_b200.template_atoms import JessTemplate, TemplateAtom

    def atom(i, mode=0, names=None, resnames=None, chain=None, resnum=None):
        return TemplateAtom(chain_id=chain if chain is not None else str(mol_1amy.column("chain_id")[i]),
                            residue_number=int(mol_1amy.column("residue_number")[i]) if resnum is None else resnum,
                            residue_names=resnames or [str(mol_1amy.column("residue_name")[i])],
                            atom_names=names or [str(mol_1amy.column("name")[i])], match_mode=mode,
                            x=float(mol_1amy.xyz[i, 0]) + 0.05 * (i % 3), y=float(mol_1amy.xyz[i, 1]),
                            z=float(mol_1amy.xyz[i, 2]) - 0.04 * (i % 2))

    resnum = mol_1amy.column("residue_number")
    idx = lambda r: np.nonzero((resnum == r) & (mol_1amy.column("chain_id") == "A"))[0]
    r87, r179, r204, r288 = idx(87), idx(179), idx(204), idx(288)
    templates = [
        JessTemplate([atom(r87[1])], id="one-atom"),
        JessTemplate([atom(r87[1]), atom(r87[5], 3), atom(r179[1]), atom(r179[2]), atom(r179[5], 3), atom(r179[6], 3),
                      atom(r204[1]), atom(r204[2])], id="groups-2-4-2"),
        JessTemplate([atom(r288[1]), atom(r288[5]), atom(r288[6], 8), atom(r204[6], 3), atom(r87[0], 1)], id="mixed-modes"),
        JessTemplate([atom(r87[1], 100), atom(r87[2], 100), atom(r179[1], 103), atom(r204[1], 100)], id="any-residue"),
    ]
    ot = oracle.OracleTemplates(templates)
    for cut, ignore in ((0.8, True), (1.5, False)):
        hits = list(pyjess.Jess(templates).query(mol_1amy, 2.0, cut, cut, max_candidates=10000, best_match=True,
                                                 ignore_chain=ignore))
        want = oracle.query([mol_1amy], ot, 2.0, cut, cut, max_candidates=10000, ignore_chain=ignore)[0]
        assert [h.template.id for h in hits] == [templates[w.template_index].id for w in want]
        assert len(hits) >= 3
        for h, w in zip(hits, want):
            if w.overflow or h.overflow:
                assert h.overflow == w.overflow
                continue
            assert h.atom_indices == w.atoms and h.rmsd == w.rmsd and h.n_complete == w.n_complete
            assert math.isnan(h.orientation) and h.device_pass


# ---- file screening path: native ingest -> device, no Molecule objects ---------------------------------

def test_scan_files_equals_matcher_run(tmp_path, active_templates):
    """``Matcher.scan_files`` (files -> packed columns natively -> GPU, chunked and prefetched) gives,
    file by file, the matches ``Matcher.run(load_molecules(files))`` gives."""
    chunk = generate_chunk(11, SynthConfig(n_residues=160), templates=active_templates, count=5)
    paths = []
    for i in range(5):
        p = tmp_path / f"synth{i}.pdb"
        p.write_text(chunk.to_pdb(i))
        paths.append(p)
    paths[2:2] = [GOLDEN / "1AMY.pdb", GOLDEN / "AF-P0DUB6-F1-model_v4.pdb"]
    for kwargs in (dict(), dict(filter_matches=False, skip_smaller_hits=True)):
        matcher = jess_run.Matcher(templates=active_templates, **kwargs)
        molecules = jess_run.load_molecules(paths)
        want = matcher.run(molecules)
        seen = 0
        for chunk_paths, ids, records in matcher.scan_files(paths, chunk_size=3, threads=2):
            assert chunk_paths == [str(p) for p in paths[seen:seen + 3]]
            for j, path in enumerate(chunk_paths):
                mol = molecules[seen + j]
                assert ids[j] == Molecule.load(path).id
                got = matcher.matches_for(mol, records[records["structure"] == j])
                expect = want.get(mol, [])
                assert [m.hit.template.id for m in got] == [m.hit.template.id for m in expect]
                assert [m.hit.atom_indices for m in got] == [m.hit.atom_indices for m in expect]
                assert [m.hit.rmsd for m in got] == [m.hit.rmsd for m in expect]
                assert [m.complete for m in got] == [m.complete for m in expect]
            seen += len(chunk_paths)
        assert seen == len(paths)
        # the same chunks pulled from a (here: single-process) ChunkQueue give the same records
        from enzymm_b200.sharding import ChunkQueue
        plain = [r for _, _, r in matcher.scan_files(paths, chunk_size=3, threads=2)]
        queued = [r for _, _, r in matcher.scan_files(paths, threads=2, queue=ChunkQueue(len(paths), 3))]
        assert len(plain) == len(queued) == 3 and all(a.tobytes() == b.tobytes() for a, b in zip(plain, queued))


def test_matcher_grows_its_hit_buffer(tmp_path, active_templates, mol_1amy, mol_af):
    """More hits than the initial buffer holds: ``run`` and ``scan_files`` enlarge it and rerun."""
    reference = jess_run.Matcher(templates=active_templates, filter_matches=False).run([mol_1amy, mol_af])
    small = jess_run.Matcher(templates=active_templates, filter_matches=False)
    small.hits_per_structure, small._hit_floor = 1, 4
    got = small.run([mol_1amy, mol_af])
    assert small.hits_per_structure > 1
    for mol in (mol_1amy, mol_af):
        assert [m.hit.template.id for m in got[mol]] == [m.hit.template.id for m in reference[mol]]
    small.hits_per_structure, small._hit_floor = 1, 4
    paths = [GOLDEN / "1AMY.pdb", GOLDEN / "AF-P0DUB6-F1-model_v4.pdb"]
    counts = [len(records) for _, _, records in small.scan_files(paths, chunk_size=1)]
    assert counts == [len(reference[mol_1amy]), len(reference[mol_af])]


def test_mixed_batch_stages_what_fits(active_templates, mol_1amy, monkeypatch):
    """One assembly too large for shared memory in a batch of ordinary structures: the ordinary ones
    are still searched from shared memory (their own launch), the large one in place -- same hits."""
    monkeypatch.setenv("EMM_STATS", "1")              # stats build: exposes staged / in-place counters
    big = generate_chunk(0, SynthConfig(n_chains=4), active_templates, 1).to_molecule(0)
    small = generate_chunk(3, SynthConfig(n_residues=200), active_templates, 3)
    mols = [small.to_molecule(0), big, mol_1amy, small.to_molecule(1), small.to_molecule(2)]
    subset = active_templates[::12]
    dist = default_distances(subset)
    eng = Engine(CompiledLibrary(subset, 2.0, dist, dist))
    try:
        compare_with_oracle(eng, subset, mols, dist)
        batch = pack_molecules(mols, eng.compiled)
        hits, stats = eng.query(batch, with_stats=True)
        assert stats["global_blobs"] >= 1 and stats["staged_bytes"] > 0
        assert stats["pairs"] == len(mols) * len(subset)
        alone = [eng.query(pack_molecules([m], eng.compiled)) for m in mols]
        for i, part in enumerate(alone):
            mine = hits[hits["structure"] == i]
            assert mine["template_index"].tolist() == part["template_index"].tolist()
            assert mine["rmsd"].tolist() == part["rmsd"].tolist()
    finally:
        eng.close()


# ---- BASELINE configs 4 and 5 at their stated parameters ---------------------------------------------

def test_config4_unfiltered_loose_cutoff_full_library(active_templates):
    """BASELINE config 4 as stated: ``--unfiltered -j 2 3.0 3.0``, ``max_candidates = 10**7``,
    the full active library x 32 synthetic 400-residue structures, against the oracle: hit set, atoms,
    RMSD bits and the number of complete assignments of every one of the ~50 k hits (up to ~10^5
    assignments superposed per pair; no pair may reach the cap)."""
    chunk = generate_chunk(0, SynthConfig(), active_templates, 32)
    mols = [chunk.to_molecule(i) for i in range(chunk.n_structures)]
    eng = Engine(CompiledLibrary(active_templates, 2.0, 3.0, 3.0))
    try:
        hits = compare_with_oracle(eng, active_templates, mols, 3.0, rmsd=2.0, max_candidates=10 ** 7)
    finally:
        eng.close()
    assert len(hits) > 1000 * len(mols) // 2 and not (hits["flags"] & HIT_OVERFLOW).any()
    assert int(hits["n_complete"].max()) > 10000
    # the same batch through the Matcher front end: --unfiltered keeps every raw hit
    params = {s: {"rmsd": 2, "distance": 3.0, "max_dynamic_distance": 3.0} for s in range(3, 9)}
    matcher = jess_run.Matcher(active_templates, jess_params=params, filter_matches=False, max_candidates=10 ** 7)
    try:
        got = matcher.run(mols[:4])
    finally:
        matcher.close()
    per_structure = np.bincount(hits["structure"], minlength=len(mols))
    assert [len(got[m]) for m in mols[:4]] == per_structure[:4].tolist()


def test_config5_masked_assemblies_skip_smaller_full_library(active_templates):
    """BASELINE config 5 as stated: 4 x 400-residue assemblies (chains A-D, ~12.5 k atoms),
    ``--conservation-cutoff 70`` applied as a real mask, ``--skip-smaller-hits``, full active library x
    16 assemblies, against the oracle's restatement of ``Matcher.run`` on explicitly masked molecules."""
    chunk = generate_chunk(0, SynthConfig(n_chains=4), active_templates, 16)
    mols = [chunk.to_molecule(i) for i in range(chunk.n_structures)]
    masked = [m.conserved(70) for m in mols]
    keeps = [np.nonzero(m.column("temperature_factor") >= 70)[0] for m in mols]
    assert all(len(m) > 9000 for m in mols) and all(0 < len(k) < len(m) for k, m in zip(keeps, mols))
    matcher = jess_run.Matcher(templates=active_templates, skip_smaller_hits=True, conservation_cutoff=70,
                               apply_conservation_mask=True)
    try:
        got = matcher.run(mols)
    finally:
        matcher.close()
    want = oracle_matcher_run(active_templates, masked, skip_smaller_hits=True, threads=16)
    assert [mols.index(k) for k in got] == list(want) and len(want) >= 4
    for mi, wm in want.items():
        gm = got[mols[mi]]
        assert [m.hit.template.id for m in gm] == [m.template.id for m in wm]
        assert [m.hit.atom_indices for m in gm] == [[int(keeps[mi][a]) for a in m.hit.atoms] for m in wm]
        assert [m.hit.rmsd for m in gm] == [m.hit.rmsd for m in wm]
        assert [m.complete for m in gm] == [m.complete for m in wm]
    # and the raw device path on the unmasked assemblies (global-memory search), every size group
    dist = default_distances(active_templates)
    eng = Engine(CompiledLibrary(active_templates, 2.0, dist, dist))
    try:
        compare_with_oracle(eng, active_templates, mols[:4], dist)
        compare_with_oracle(eng, active_templates, mols[:4], dist, cutoff=70.0, oracle_molecules=masked[:4],
                            atom_maps=keeps[:4])
    finally:
        eng.close()


# ---- structures beyond 65 535 atoms; structures the device refuses -----------------------------------

def _mega_molecule(templates, n_assemblies=8):
    """One structure of ~100 k atoms: ``n_assemblies`` 4-chain synthetic assemblies on a 95 A lattice,
    every chain with its own two-letter id."""
    from enzymm_b200.structures import _COLUMNS
    chunk = generate_chunk(0, SynthConfig(n_chains=4), templates, n_assemblies)
    parts = [chunk.to_molecule(i) for i in range(n_assemblies)]
    cols = {k: np.concatenate([m.column(k) for m in parts]) for k, _ in _COLUMNS}
    xyz = np.concatenate([m.xyz + 95.0 * np.array([i & 1, (i >> 1) & 1, (i >> 2) & 1]) for i, m in enumerate(parts)])
    cols["chain_id"] = np.concatenate([np.char.add(m.column("chain_id").astype("U1"), "ABCDEFGH"[i])
                                       for i, m in enumerate(parts)]).astype("U2")
    cols["serial"] = (np.arange(len(xyz)) % 99999 + 1).astype(np.int32)
    return Molecule._from_columns(cols, xyz, "mega")


def test_structure_beyond_65535_atoms(active_templates, mol_1amy):
    """A 100 k-atom assembly (32-bit index arrays in its blob, searched in place) next to an ordinary
    structure in the same batch: hit set, atoms, RMSD bits and counts equal the oracle's; the reference
    has no size limit, round 1 refused anything above 65 535 kept atoms."""
    subset = active_templates[::40]
    mega = _mega_molecule(subset)
    assert len(mega) > 100000
    dist = default_distances(subset)
    eng = Engine(CompiledLibrary(subset, 2.0, dist, dist))
    try:
        hits = compare_with_oracle(eng, subset, [mol_1amy, mega], dist)
        assert (hits["structure"] == 1).sum() >= 3 and int(hits["atoms"].max()) > 65535
        cells = eng.query(pack_molecules([mega], eng.compiled), cell_threshold=64)      # wide cell list
        plain = hits[hits["structure"] == 1]
        assert np.array_equal(cells["atoms"], plain["atoms"]) and np.array_equal(cells["rmsd"], plain["rmsd"])
    finally:
        eng.close()
    got = jess_run.Matcher(templates=subset).run([mega, mol_1amy])
    want = oracle_matcher_run(subset, [mega, mol_1amy], threads=2)
    assert {k.id: [m.hit.atom_indices for m in v] for k, v in got.items()} == \
           {[mega, mol_1amy][i].id: [m.hit.atoms for m in v] for i, v in want.items()}


def test_refused_structure_does_not_cost_the_batch(full_engine, active_templates, mol_1amy, mol_af):
    """one structure outside the input contract used to fail the whole batch.  Now it is
    skipped and named; every other structure's hits are delivered (engine: on the exception; Matcher:
    a warning and the results)."""
    good = pack_molecules([mol_1amy, mol_af], full_engine.compiled)
    n1 = len(mol_1amy)
    # middle structure: residue ordinals that decrease (status 1)
    atom_off = np.array([0, n1, n1 + 4, n1 + 4 + len(mol_af)])
    xyz = np.concatenate([good.xyz[:n1], np.zeros((4, 3)), good.xyz[n1:]])
    klass = np.concatenate([good.klass[:n1], np.ones(4, dtype=np.uint16), good.klass[n1:]])
    residue = np.concatenate([good.residue[:n1], np.array([0, 1, 0, 1], dtype=np.int32), good.residue[n1:]])
    batch = PackedBatch(atom_off, xyz, klass, residue)
    with pytest.raises(EngineError) as info:
        full_engine.query(batch)
    exc = info.value
    assert exc.status == -5 and exc.bad_structures == {1: 1}
    clean = full_engine.query(good)
    relabel = exc.hits.copy()
    relabel["structure"] = np.where(relabel["structure"] == 2, 1, relabel["structure"])
    assert len(clean) == 13 + 11 and relabel.tobytes() == clean.tobytes()
    # a residue with more than 1023 kept atoms (status 3)
    n = 1100
    blob = PackedBatch(np.array([0, n]), np.random.default_rng(0).normal(size=(n, 3)) * 20,
                       np.full(n, good.klass[good.klass > 0][0], dtype=np.uint16), np.zeros(n, dtype=np.int32))
    with pytest.raises(EngineError) as info:
        full_engine.query(blob)
    assert info.value.bad_structures == {0: 3} and len(info.value.hits) == 0


def test_pair_splitting_changes_nothing(full_engine, active_templates, mol_1amy, mol_af):
    """One (template, structure) pair may be split over the warps of its CTA (subtrees donated to idle
    warps, merged by minimum RMSD / lexicographic tie-break / summed counts).  ``donate_after=1`` makes
    every pair donate whenever a warp is idle -- far more splitting than the default -- and nothing in
    the results may move: against the oracle, and bit for bit against the unsplit search."""
    dist = default_distances(active_templates)
    split = compare_with_oracle(full_engine, active_templates, [mol_1amy, mol_af], dist, donate_after=1)
    plain = full_engine.query(pack_molecules([mol_1amy, mol_af], full_engine.compiled), donate_after=-1)
    assert split.tobytes() == plain.tobytes() and len(plain) == 24
    # few templates: most warps idle from the start; loose cutoff: deep, wide trees; exhaustive counts
    subset = [t for t in active_templates if t.effective_size in (3, 4)][::150] + \
             [t for t in active_templates if t.effective_size >= 6][::40]
    chunk = generate_chunk(2, SynthConfig(), active_templates, 6)
    mols = [chunk.to_molecule(i) for i in range(chunk.n_structures)] + [mol_1amy]
    eng = Engine(CompiledLibrary(subset, 2.0, 3.0, 3.0))
    try:
        for after in (1, 7):
            hits = compare_with_oracle(eng, subset, mols, 3.0, max_candidates=10 ** 7, donate_after=after)
            assert len(hits) > 20 and int(hits["n_complete"].max()) > 1000
        unsplit = eng.query(pack_molecules(mols, eng.compiled), max_candidates=10 ** 7, donate_after=-1)
        assert hits.tobytes() == unsplit.tobytes()
        capped = compare_with_oracle(eng, subset, mols, 3.0, max_candidates=50, donate_after=1)
        assert (capped["flags"] & HIT_OVERFLOW).any()
    finally:
        eng.close()
    # a 4-chain assembly searched in place, split
    big = generate_chunk(0, SynthConfig(n_chains=4), active_templates, 1).to_molecule(0)
    few = active_templates[::300]
    eng = Engine(CompiledLibrary(few, 2.0, default_distances(few), default_distances(few)))
    try:
        compare_with_oracle(eng, few, [big], default_distances(few), donate_after=1)
    finally:
        eng.close()


def test_scan_files_over_several_devices_from_one_process(tmp_path, active_templates):
    """The whole-box call: ``scan_files(paths, devices=[...])`` runs one worker per listed GPU in this
    process, chunks handed out dynamically, results back in input order.  On a one-GPU box the two
    workers share device 0 -- the threading, hand-out and re-ordering are what is under test; the
    records must equal the single-worker scan, and the table written through it must not change."""
    import io
    chunk = generate_chunk(4, SynthConfig(n_residues=150), active_templates, 40)
    paths = []
    for i in range(chunk.n_structures):
        path = tmp_path / f"s{i:03d}.pdb"
        path.write_text(chunk.to_pdb(i))
        paths.append(path)
    matcher = jess_run.Matcher(active_templates)
    try:
        single = list(matcher.scan_files(paths, chunk_size=7))
        multi = list(matcher.scan_files(paths, chunk_size=7, devices=[0, 0]))
        assert [c for c, _, _ in multi] == [c for c, _, _ in single]
        assert all(a[2].tobytes() == b[2].tobytes() for a, b in zip(single, multi)) and sum(len(r) for _, _, r in multi) > 10
        one, two = io.StringIO(), io.StringIO()
        rows = matcher.scan_to_tsv(paths, one, chunk_size=7)
        assert matcher.scan_to_tsv(paths, two, chunk_size=7, devices=[0, 0]) == rows > 5
        assert one.getvalue() == two.getvalue()
        # and the table equals the reference-shaped route: Match objects, one dump per match, per chunk
        want = io.StringIO()
        first = True
        for lo in range(0, len(paths), 7):
            molecules = jess_run.load_molecules(paths[lo:lo + 7])
            for matches in matcher.run(molecules).values():
                for j, match in enumerate(matches):
                    match.index = j + 1
                    match.dump(want, header=first)
                    first = False
        assert one.getvalue() == want.getvalue()
    finally:
        matcher.close()


def test_staging_policy_keeps_the_l1_carveout(active_templates, monkeypatch):
    """A staged blob beyond ~97 KB pushes the CTA past 196 KB of shared memory and the SM's L1 from 60
    to 28 KB for the whole launch.  A lone structure of that size in a batch of ordinary ones is
    therefore searched in place; when such structures are more than a tenth of the batch they are
    staged after all.  Hits never depend on the route."""
    monkeypatch.setenv("EMM_STATS", "1")
    small = generate_chunk(5, SynthConfig(n_residues=200), active_templates, 20)
    mid = generate_chunk(6, SynthConfig(n_residues=560), active_templates, 4)      # ~110 KB staged with the full library
    dist = default_distances(active_templates)
    eng = Engine(CompiledLibrary(active_templates, 2.0, dist, dist))
    try:
        lone = [small.to_molecule(i) for i in range(20)] + [mid.to_molecule(0)]
        many = [small.to_molecule(i) for i in range(6)] + [mid.to_molecule(i) for i in range(4)]
        for mols, in_place in ((lone, 1), (many, 0)):
            hits, stats = eng.query(pack_molecules(mols, eng.compiled), with_stats=True)
            assert (stats["global_blobs"] > 0) == bool(in_place), stats
            alone = [eng.query(pack_molecules([m], eng.compiled)) for m in mols]
            for i, part in enumerate(alone):
                mine = hits[hits["structure"] == i]
                assert mine["template_index"].tolist() == part["template_index"].tolist()
                assert mine["rmsd"].tolist() == part["rmsd"].tolist() and np.array_equal(mine["atoms"], part["atoms"])
        compare_with_oracle(eng, active_templates, lone[-2:], dist)
    finally:
        eng.close()


def test_download_into_a_caller_buffer(full_engine, mol_1amy, mol_af):
    """``Session.download(out=...)``: hits written straight into a slice of the caller's (pinned) buffer,
    back to back over several batches -- what the strong-scaling merge is built on -- equal the copies
    ``download()`` returns; a slice that is too small is the loud ``EMM_ERR_CAPACITY``."""
    from enzymm_b200.engine import HIT_DTYPE
    batches = [pack_molecules([m], full_engine.compiled) for m in (mol_1amy, mol_af, mol_1amy)]
    room = np.zeros(64, dtype=HIT_DTYPE)
    used, want = 0, []
    sess = full_engine.session_for(max(b.n_atoms for b in batches), 1)
    for batch in batches:
        sess.upload(batch)
        sess.run()
        want.append(sess.download())
        got = sess.download(out=room[used:])
        assert got.base is not None and got.tobytes() == want[-1].tobytes()
        used += len(got)
    assert used == 13 + 11 + 13 and room[:used].tobytes() == np.concatenate(want).tobytes()
    with pytest.raises(EngineError) as info:
        sess.download(out=room[:5])
    assert info.value.status == -4
    with pytest.raises(ValueError):
        sess.download(out=np.zeros(8, dtype=np.int32))
