"""An independent check of the oracle's search: a plain Python enumeration written from the rule
set (typing by ``library.type_match``, same-residue rule, injectivity, every pairwise
distance within the cutoff, Kabsch RMSD by SVD, best = minimum RMSD) must find the same number of
complete assignments, the same best RMSD and the same atoms as ``oracle/jess_oracle.c``.  The
golden vectors pin the oracle at a handful of points; this pins its completeness on small inputs."""
import numpy as np

import oracle
from enzymm_b200.library import type_match
from enzymm_b200.synth import SynthConfig, generate_chunk
from enzymm_b200.templates import load_templates
from helpers import svd_kabsch


def brute_force(template, mol, rmsd_threshold, cutoff, max_dynamic=None, ignore_chain=True):
    atoms = list(template)
    m = len(atoms)
    t_xyz = np.array([(a.x, a.y, a.z) for a in atoms])
    names, resnames = mol.column("name"), mol.column("residue_name")
    q_res = list(zip(mol.column("chain_id").tolist(), mol.column("residue_number").tolist()))
    cands = [[i for i in range(len(mol)) if type_match(a.match_mode, tuple(a.residue_names), tuple(a.atom_names),
                                                        str(resnames[i]), str(names[i]))] for a in atoms]
    t_dist = np.linalg.norm(t_xyz[:, None] - t_xyz[None], axis=2)
    same_res = [[(a.chain_id, a.residue_number) == (b.chain_id, b.residue_number) for b in atoms] for a in atoms]
    # tolerance of pair (i, j): the cutoff, widened by both atoms' distance weights up to max_dynamic
    if max_dynamic is None or max_dynamic == cutoff:
        tol = [[cutoff] * m for _ in range(m)]
    else:
        tol = [[min(cutoff + a.distance_weight + b.distance_weight, max_dynamic) for b in atoms] for a in atoms]
    q_chain = mol.column("chain_id").tolist()
    xyz = mol.xyz
    best = (None, None)
    count = 0
    assign = []

    def place(k):
        nonlocal best, count
        if k == m:
            if not ignore_chain:      # template atoms share a chain exactly when their query atoms do
                for i in range(m):
                    for j in range(i + 1, m):
                        if (atoms[i].chain_id == atoms[j].chain_id) != (q_chain[assign[i]] == q_chain[assign[j]]):
                            return
            count += 1
            rmsd, _ = svd_kabsch(t_xyz, xyz[assign])
            if rmsd <= rmsd_threshold and (best[0] is None or rmsd < best[0] - 1e-12):
                best = (rmsd, list(assign))
            return
        for c in cands[k]:
            if c in assign:
                continue
            ok = True
            for j in range(k):
                if same_res[k][j] and q_res[c] != q_res[assign[j]]:
                    ok = False
                    break
                if abs(np.linalg.norm(xyz[c] - xyz[assign[j]]) - t_dist[k, j]) > tol[k][j]:
                    ok = False
                    break
            if ok:
                assign.append(c)
                place(k + 1)
                assign.pop()

    place(0)
    return count, best


def test_oracle_equals_brute_force_on_small_inputs():
    templates = list(load_templates(subset="3_residues/results/csa3d_00"))[:60]
    chunk = generate_chunk(9, SynthConfig(n_residues=45, max_motifs=2), templates, 4)
    mols = [chunk.to_molecule(i) for i in range(4)]
    for cutoff in (0.9, 2.0):
        # max_dynamic_distance == distance_cutoff: per-atom distance weights cannot widen any tolerance
        raw = oracle.query_raw(mols, oracle.OracleTemplates(templates), 2.0, cutoff, cutoff, max_candidates=10 ** 9,
                               ignore_chain=True, threads=4)
        found = complete = 0
        for mi, mol in enumerate(mols):
            for ti, t in enumerate(templates):
                count, (rmsd, atoms) = brute_force(t, mol, 2.0, cutoff)
                r = raw[mi, ti]
                assert int(r["n_complete"]) == count, (mi, ti)
                assert bool(r["found"]) == (rmsd is not None), (mi, ti)
                complete += count
                if rmsd is not None:
                    found += 1
                    assert abs(float(r["rmsd"]) - rmsd) < 1e-9, (mi, ti)
                    assert r["atoms"][:len(atoms)].tolist() == atoms, (mi, ti)
        assert found > 0 and complete > found        # the comparison is not vacuous


def test_oracle_equals_brute_force_dynamic_distances_and_chain_rule():
    """The two rule variants no reference vector pins: distance weights widening the
    tolerance up to ``max_dynamic_distance``, and ``ignore_chain=False`` on a two-chain structure."""
    templates = [t for t in load_templates(subset="3_residues/results/csa3d_00") if any(a.distance_weight for a in t)][:40]
    chunk = generate_chunk(4, SynthConfig(n_residues=24, n_chains=2, max_motifs=2), templates, 3)
    mols = [chunk.to_molecule(i) for i in range(3)]
    found = {}
    completes = {}
    for cutoff, dyn, ignore_chain in ((0.9, 2.5, True), (2.0, 2.0, True), (2.0, 2.0, False), (1.0, 2.5, False)):
        raw = oracle.query_raw(mols, oracle.OracleTemplates(templates), 2.0, cutoff, dyn, max_candidates=10 ** 9,
                               ignore_chain=ignore_chain, threads=4)
        key = (cutoff, dyn, ignore_chain)
        found[key] = completes[key] = 0
        for mi, mol in enumerate(mols):
            for ti, t in enumerate(templates):
                count, (rmsd, atoms) = brute_force(t, mol, 2.0, cutoff, dyn, ignore_chain)
                r = raw[mi, ti]
                assert int(r["n_complete"]) == count, (key, mi, ti)
                assert bool(r["found"]) == (rmsd is not None)
                completes[key] += count
                if rmsd is not None:
                    found[key] += 1
                    assert abs(float(r["rmsd"]) - rmsd) < 1e-9 and r["atoms"][:len(atoms)].tolist() == atoms
    assert found[(0.9, 2.5, True)] > 0                                           # dynamic distances exercised
    assert 0 < completes[(2.0, 2.0, False)] < completes[(2.0, 2.0, True)]        # the chain rule rejects some, not all
