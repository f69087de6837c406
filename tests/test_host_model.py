"""Host-side logic: PDB ingest, template parsing, typing, search plans, packing (no GPU)."""
import math

import numpy as np
import pytest

from conftest import GOLDEN
from enzymm_b200 import pyjess
from enzymm_b200.library import CompiledLibrary, type_match
from enzymm_b200.packing import chain_codes, pack_molecules, residue_ordinals
from enzymm_b200.structures import Atom, Molecule
from enzymm_b200.template_atoms import JessTemplate, TemplateAtom
from enzymm_b200.templates import (Cluster, Residue, Template, Vec3, check_template, load_templates,
                                   rank_order)


def test_molecule_load(mol_1amy, mol_af):
    assert len(mol_1amy) == 3339 and bool(mol_1amy)
    assert (mol_1amy.column("name") == "CA").sum() > 400          # CA atoms + 3 calcium HETATM
    a = mol_1amy.atom(0)
    assert (a.name, a.residue_name, a.chain_id, a.residue_number) == ("N", "GLN", "A", 1)
    assert mol_af.id is None and len(mol_af) == 4079
    assert len(mol_af.conserved(80)) == 3933
    assert mol_af.conserved(0) == mol_af and hash(mol_af.conserved(0)) == hash(mol_af)
    assert not Molecule()
    with pytest.raises(FileNotFoundError):
        Molecule.load("/no/such/file.pdb")
    with pytest.raises(IsADirectoryError):
        Molecule.load(str(GOLDEN))


def test_molecule_endmdl():
    text = ("ATOM      1  N   GLY A   1       0.000   0.000   0.000  1.00 10.00           N\n"
            "ENDMDL\n"
            "ATOM      2  CA  GLY A   1       1.000   0.000   0.000  1.00 10.00           C\n")
    assert len(Molecule.loads(text)) == 1


def test_template_atom_loads():
    a = TemplateAtom.loads("ATOM      3  CG ZASP A 262      49.175  39.646  17.664 DE    1.90 ")
    assert a.match_mode == 3 and a.atom_names == ("CG",) and a.residue_names == ("ASP", "GLU")
    assert (a.chain_id, a.residue_number, a.distance_weight) == ("A", 262, 1.9)
    assert (a.x, a.y, a.z) == (49.175, 39.646, 17.664)
    b = TemplateAtom.loads("ATOM      0  CG ZHISAA 180      17.497  30.652  21.394 H     0.55 ")
    assert b.chain_id == "AA" and b.residue_names == ("HIS",) and b.residue_number == 180
    c = TemplateAtom.loads("ATOM      0  CA ZANY A  70       9.165   4.861  36.502 AXSCG 0.40 ")
    assert c.residue_names == ("ANY", "ALA", "XXX", "SER", "CYS", "GLY") and c.distance_weight == 0.4
    assert a == a.copy() and hash(a) == hash(a.copy()) and a != b
    with pytest.raises(ValueError):
        TemplateAtom.loads("ATOM      1  N   GLN A   1      -8.553  67.654  28.389  1.00 31.38           N")
    with pytest.raises(ValueError):
        TemplateAtom.loads("HETATM    1  N   GLN A   1      -8.553  67.654  28.389")


def test_library_loads(all_templates):
    import collections
    assert len(all_templates) == 7607                                   # test_template.py:35
    sizes = collections.Counter(t.effective_size for t in all_templates)
    assert [sizes[s] for s in (8, 7, 6, 5, 4, 3)] == [179, 359, 654, 1046, 1841, 2701]
    assert sum(len(t) for t in all_templates if t.effective_size >= 3) == 90543


def test_template_metadata(template_1uh3):
    t = template_1uh3
    assert (t.pdb_id, t.mcsa_id, t.uniprot_id) == ("1uh3", 285, "Q60053")
    assert t.cluster == Cluster(1, 1, 1) and t.dimension == 5 and t.effective_size == 5
    assert t.ec == ("3.2.1.10", "3.2.1.135") and t.cath == ("2.60.40.10", "2.60.40.1180", "3.20.20.80")
    assert not t.multimeric and t.relative_order == rank_order([396, 262, 356, 471, 472])
    assert [r.orientation_vector_indices for r in t.residues] == [(0, 9), (0, 9), (0, 9), (0, 1), (0, 9)]
    assert t == t.copy() and hash(t) == hash(t.copy())
    assert check_template(t, warn=False)


def test_bad_templates():
    bad = GOLDEN / "bad_templates"
    for name in ("hetatm.pdb", "malformed_residues_1.pdb", "malformed_residues_2.pdb", "malformed_residues_3.pdb",
                 "not_a_template.pdb", "bad_atom_names_1.pdb", "bad_atom_names_2.pdb", "bad_atom_names_3.pdb"):
        with pytest.raises(ValueError):
            Template.loads((bad / name).read_text(), warn=True)
    with pytest.raises(KeyError):
        Template.loads((bad / "unk_residue.pdb").read_text(), warn=True)
    Template.loads((bad / "no_remark.pdb").read_text())
    Template.loads((bad / "new_remark.pdb").read_text())
    with pytest.raises(NotADirectoryError):
        list(load_templates("/some/bogus/folder"))
    with pytest.raises(NotADirectoryError):
        list(load_templates(bad / "pdb_id_none.pdb"))       # a file, not a folder (test_template.py:23-28)
    with pytest.raises(IndexError):
        Template.loads((bad / "missing_cluster_annotation.pdb").read_text(), warn=True)
    assert Template.loads((bad / "pdb_id_none.pdb").read_text(), warn=True).pdb_id is None
    with pytest.raises(ValueError):
        list(load_templates(GOLDEN))           # a folder of non-template .pdb files


def test_vec3():
    v = Vec3(1.0, 2.0, 3.0)
    assert v.norm == math.sqrt(14) and v.normalize().x == 1 / math.sqrt(14)
    assert Vec3(0, 0, 0).normalize() == Vec3(0, 0, 0)
    assert (v + 5) == Vec3(6, 7, 8) and (v - Vec3(1, 1, 1)) == Vec3(0, 1, 2) and (v / 2) == Vec3(0.5, 1, 1.5)
    assert v @ Vec3(1, 0, 0) == 1.0
    assert Vec3(1.00000000001, 0, 0).angle_to(Vec3(-1, 0, 0)) == pytest.approx(math.pi)
    with pytest.raises(TypeError):
        v + (0, 3, 5)
    with pytest.raises(ValueError):
        Vec3(float("nan"), 0, 0)


def test_type_match_rules():
    assert type_match(0, ("HIS",), ("CG",), "HIS", "CG") and not type_match(0, ("HIS",), ("CG",), "HIS", "CB")
    assert not type_match(0, ("HIS",), ("CG",), "ASP", "CG")
    assert type_match(3, ("ASP", "GLU"), ("OD1",), "GLU", "OE2") and not type_match(3, ("ASP",), ("OD1",), "ASP", "CG")
    assert type_match(8, ("HIS",), ("ND1",), "HIS", "CD2") and not type_match(8, ("HIS",), ("ND1",), "HIS", "NE2")
    assert type_match(1, ("ASN", "GLN"), ("OD1",), "GLN", "NE2") and not type_match(1, ("ASN",), ("OD1",), "ASN", "CG")
    assert type_match(100, ("ANY",), ("CA",), "TRP", "CA") and not type_match(100, ("ANY",), ("CA",), "TRP", "CB")
    with pytest.raises(ValueError):
        type_match(7, ("ALA",), ("CA",), "ALA", "CA")


def test_compiled_library_invariants(active_templates, mol_1amy):
    subset = active_templates[::40]
    lib = CompiledLibrary(subset, 2.0, 1.5, 1.5)
    cm = lib.compat_matrix()
    assert cm.shape == (lib.n_ttype, lib.class_words) and not cm[:, 0].astype(np.uint32)[0] & 1
    for ti, t in enumerate(subset):
        lo, hi = lib.atom_off[ti], lib.atom_off[ti + 1]
        assert sorted(lib.plan_atom[lo:hi].tolist()) == list(range(hi - lo))
        assert lib.plan_src[lo] < 0
        atoms = list(t)
        for k in range(hi - lo):
            src = int(lib.plan_src[lo + k])
            a = atoms[lib.plan_atom[lo + k]]
            if src >= 0:      # follower: same template residue as its leader, which is itself a leader
                lead = atoms[lib.plan_atom[lo + src]]
                assert (lead.chain_id, lead.residue_number) == (a.chain_id, a.residue_number)
                assert lib.plan_src[lo + src] < 0
            assert lib.keys[lib.plan_ttype[lo + k]] == a.typing_key()
    # typing matrix agrees with the predicate for every atom of a real structure
    klass = lib.classify(mol_1amy.column("residue_name"), mol_1amy.column("name"))
    cm = lib.compat_matrix()
    for tt in range(0, lib.n_ttype, 7):
        mode, resnames, names = lib.keys[tt]
        want = np.array([type_match(mode, resnames, names, str(r), str(n))
                         for r, n in zip(mol_1amy.column("residue_name"), mol_1amy.column("name"))])
        got = ((cm[tt][klass >> 5] >> (klass & 31)) & 1).astype(bool)
        assert np.array_equal(got, want)
    assert (klass[mol_1amy.column("residue_name") == "HOH"] == 0).all() or lib.n_classes > 0


def test_residue_ordinals_and_packing(mol_1amy, active_templates):
    lib = CompiledLibrary(active_templates[:20], 2.0, 1.5, 1.5)
    batch = pack_molecules([mol_1amy, Molecule(), mol_1amy.conserved(20)], lib)
    assert batch.n_structures == 3 and batch.atom_off.tolist()[1:3] == [3339, 3339]
    assert batch.atom_id is None
    for s in range(3):
        r = batch.residue[batch.atom_off[s]:batch.atom_off[s + 1]]
        assert (np.diff(r) >= 0).all()
    # a residue split over two runs gets reordered, and atom_id maps back
    chain = chain_codes(np.array(["A", "A", "A", "A"]))
    ordinal, order = residue_ordinals(chain, np.array([5, 6, 5, 7]))
    assert order.tolist() == [0, 2, 1, 3] and ordinal.tolist() == [0, 0, 1, 2]
    ordinal, order = residue_ordinals(chain, np.array([5, 5, 6, 7]))
    assert order is None and ordinal.tolist() == [0, 0, 1, 2]


def test_pyjess_surface():
    assert isinstance(pyjess.__version__, str)
    for name in ("Atom", "Molecule", "TemplateAtom", "Template", "Jess", "Query", "Hit"):
        assert hasattr(pyjess, name)
    atom = TemplateAtom(chain_id="A", residue_number=51, residue_names=["ANY"], atom_names=["C"],
                        distance_weight=1.5, match_mode=1, x=1.0, y=0, z=-99.5)
    t = JessTemplate([atom], id="x")
    assert t.dimension == 1 and len(t) == 1 and list(t) == [atom] and t.copy() == t
    with pytest.raises(NotImplementedError):
        pyjess.Jess([t]).query(Molecule(), 2, 1, 1)          # best_match=False is not implemented


def test_native_pdb_ingest_matches_python_parser(mol_1amy):
    """emm_pdb.cpp (C ABI, host only) yields the same columns and bit-identical doubles as the
    pure-Python fixed-column reader; the threaded batch loader returns the same molecules."""
    from enzymm_b200.structures import _COLUMNS, _parse_pdb_text, load_many
    for name in ("1AMY.pdb", "AF-P0DUB6-F1-model_v4.pdb", "1AMY_matches_query_included.pdb"):
        native = Molecule.load(GOLDEN / name)
        cols, xyz, header = _parse_pdb_text(open(GOLDEN / name))
        assert np.array_equal(native.xyz, xyz) and native.id == header
        for key, _ in _COLUMNS:
            assert np.array_equal(native.column(key), cols[key]), key
    rng = np.random.default_rng(5)
    values = [float(f"{v:8.3f}") for v in rng.uniform(-999, 9999, 5000)]
    text = "".join(f"ATOM  {i % 99999:>5}  CA  GLY A{i % 9999:>4}    {v:8.3f}{v:8.3f}{v:8.3f}  1.00 50.00           C\n"
                   for i, v in enumerate(values))
    assert np.array_equal(Molecule.loads(text).xyz[:, 2], np.array(values))
    many = load_many([GOLDEN / "1AMY.pdb", GOLDEN / "AF-P0DUB6-F1-model_v4.pdb"] * 3, threads=3)
    assert len(many) == 6 and many[0] == mol_1amy and many[2] == mol_1amy and many[1].id is None
    assert load_many([GOLDEN / "1AMY.pdb"], ids=["x"])[0].id == "x"
    with pytest.raises(FileNotFoundError):
        load_many(["/no/such.pdb"])
    with pytest.raises(ValueError):
        Molecule.loads("ATOM      1  CA  GLY A   1      xx.xxx   0.000   0.000\n")


def test_compiled_library_cache(tmp_path, active_templates):
    subset = active_templates[::50]
    dist = [0.9 if t.effective_size == 3 else 2.0 for t in subset]
    first = CompiledLibrary.cached(subset, 2.0, dist, dist, tmp_path)
    again = CompiledLibrary.cached(subset, 2.0, dist, dist, tmp_path)
    assert len(list(tmp_path.glob("emm_library_*.pkl"))) == 1
    for name in ("atom_off", "plan_atom", "plan_ttype", "plan_src", "plan_anchor", "pair_dist", "lr_index", "lr_table"):
        assert np.array_equal(getattr(first, name), getattr(again, name)), name
    assert np.array_equal(first.compat_matrix(), again.compat_matrix())
    assert again.templates[0] is subset[0]
    CompiledLibrary.cached(subset, 2.0, 1.5, 1.5, tmp_path)             # other thresholds: a second entry
    assert len(list(tmp_path.glob("emm_library_*.pkl"))) == 2
    # a cache file is data: one that names a callable is refused instead of executed
    import pickle
    from enzymm_b200.library import _TablesUnpickler
    evil = tmp_path / "evil.pkl"
    evil.write_bytes(pickle.dumps({"x": print}))
    with pytest.raises(pickle.UnpicklingError):
        _TablesUnpickler(evil.open("rb")).load()
    for path in tmp_path.glob("emm_library_*.pkl"):
        assert (path.stat().st_mode & 0o077) == 0


def test_load_molecules_ids_and_errors(tmp_path):
    from enzymm_b200.matcher import load_molecules
    a = tmp_path / "a" / "1AMY.pdb"
    b = tmp_path / "b" / "1AMY.pdb"
    for p in (a, b):
        p.parent.mkdir()
        p.write_text((GOLDEN / "1AMY.pdb").read_text())
    empty = tmp_path / "empty.pdb"
    empty.write_text("REMARK nothing\n")
    with pytest.warns(UserWarning):
        mols = load_molecules([a, b, empty], conservation_cutoff=80, warn=True)
    assert [m.id for m in mols] == ["1AMY", "1AMY_2"] and len(mols[0]) == 3339     # cutoff result is discarded upstream
    with pytest.raises(FileNotFoundError):
        load_molecules([tmp_path / "missing.pdb"])
    with pytest.raises(IsADirectoryError):
        load_molecules([tmp_path])


def test_packed_batch_binary_cache(tmp_path, active_templates, mol_1amy):
    from enzymm_b200.engine import PackedBatch
    lib = CompiledLibrary(active_templates[:30], 2.0, 1.5, 1.5)
    batch = pack_molecules([mol_1amy, mol_1amy.conserved(30)], lib)
    batch.save(tmp_path / "batch.npz")
    again = PackedBatch.load(tmp_path / "batch.npz")
    for name in ("atom_off", "xyz", "klass", "residue", "bfactor", "chain"):
        assert np.array_equal(getattr(batch, name), getattr(again, name)), name
    assert again.atom_id is None and again.n_structures == 2


def _assert_batches_equal(a, b):
    for name in ("atom_off", "xyz", "klass", "residue", "bfactor", "chain", "atom_id"):
        x, y = getattr(a, name), getattr(b, name)
        assert (x is None) == (y is None), name
        if x is not None:
            assert x.dtype == y.dtype and np.array_equal(x, y), name


def test_pack_files_matches_molecule_path(tmp_path, active_templates):
    """files -> PackedBatch natively == pack_molecules(load_many(files)), incl. a split residue,
    an empty file, CRLF line ends and a multi-model file (first model only)."""
    from enzymm_b200.packing import pack_files
    from enzymm_b200.structures import load_many
    lib = CompiledLibrary(active_templates[:40], 2.0, 1.5, 1.5)
    text = (GOLDEN / "1AMY.pdb").read_text()
    lines = [l for l in text.splitlines() if l.startswith(("ATOM", "HETATM"))]
    split = tmp_path / "split.pdb"          # residue 1's last atoms reappear after residue 3
    first = [l for l in lines if int(l[22:26]) == 1]
    rest = [l for l in lines if 1 < int(l[22:26]) <= 3]
    split.write_text("\n".join(first[:3] + rest + first[3:] + lines[40:200]) + "\n")
    empty = tmp_path / "empty.pdb"
    empty.write_text("REMARK nothing here\nEND\n")
    crlf = tmp_path / "crlf.pdb"
    crlf.write_bytes(("\r\n".join(["HEADER    HYDROLASE" + " " * 43 + "9XYZ"] + lines[:300]) + "\r\n").encode())
    models = tmp_path / "models.pdb"
    models.write_text("MODEL        1\n" + "\n".join(lines[:50]) + "\nENDMDL\nMODEL        2\n" + "\n".join(lines[50:90]) + "\nENDMDL\n")
    paths = [GOLDEN / "1AMY.pdb", split, empty, GOLDEN / "AF-P0DUB6-F1-model_v4.pdb", crlf, models]
    for subset in (paths, [paths[0], paths[3]], [empty], []):
        expect = pack_molecules(load_many(subset), lib)
        for threads in (1, 3):
            got, ids = pack_files(subset, lib, threads=threads)
            _assert_batches_equal(expect, got)
            assert ids == [m.id for m in load_many(subset)]
    got, _ = pack_files(paths, lib)
    assert got.atom_id is not None and got.n_structures == 6
    assert np.diff(got.atom_off).tolist()[2] == 0 and np.diff(got.atom_off).tolist()[5] == 50
    no_chain, _ = pack_files(paths[:1], lib, with_chain=False)
    assert no_chain.chain is None


def test_pack_files_errors(tmp_path, active_templates):
    from enzymm_b200.packing import pack_files
    lib = CompiledLibrary(active_templates[:5], 2.0, 1.5, 1.5)
    with pytest.raises(FileNotFoundError):
        pack_files([tmp_path / "missing.pdb"], lib)
    with pytest.raises(IsADirectoryError):
        pack_files([tmp_path], lib)
    bad = tmp_path / "bad.pdb"
    bad.write_text("ATOM      1  N   ALA A   1      xx.xxx  22.000  33.000  1.00  0.00           N\n")
    with pytest.raises(ValueError, match="malformed PDB coordinate record"):
        pack_files([bad], lib)


def test_pack_files_coordinates_are_python_floats(tmp_path, active_templates):
    """The fixed-layout number path of the native packer yields exactly float(text)."""
    from enzymm_b200.packing import pack_files
    lib = CompiledLibrary(active_templates[:5], 2.0, 1.5, 1.5)
    rng = np.random.default_rng(5)
    texts = ["%8.3f" % v for v in rng.uniform(-999.0, 9999.0, 3000)]
    texts += ["  -0.000", "   0.000", "9999.999", "-999.999", "   -.500", "    .250", "  1e+002", " +12.500", "12.5    ", "     7.0"]
    while len(texts) % 3:
        texts.append("   1.000")
    lines = []
    for i in range(0, len(texts), 3):
        x, y, z = texts[i:i + 3]
        occ, bf = ("%6.2f" % rng.uniform(0, 1), "%6.2f" % rng.uniform(-9, 99)) if i % 2 else ("  1.0 ", "")
        lines.append(f"ATOM  {i % 99999:5d}  CA  ALA A{i % 9999:4d}    {x}{y}{z}{occ}{bf}".ljust(80 if i % 4 else 60))
    path = tmp_path / "numbers.pdb"
    path.write_text("\n".join(lines) + "\n")
    batch, _ = pack_files([path], lib)
    want = np.asarray([float(t) for t in texts]).reshape(-1, 3)
    assert batch.xyz.tobytes() == want.tobytes()
    want_bf = np.asarray([float(l[60:66]) if l[60:66].strip() else 0.0 for l in lines]).astype(np.float32)
    assert batch.bfactor.tobytes() == want_bf.tobytes()
    from enzymm_b200.structures import load_many
    (mol,) = load_many([path])                       # the Molecule reader shares the number paths
    assert mol.xyz.tobytes() == want.tobytes()
    assert mol.column("temperature_factor").astype(np.float32).tobytes() == want_bf.tobytes()
    assert Molecule.load(path).xyz.tobytes() == want.tobytes()


def test_leader_order_is_a_valid_plan(active_templates):
    """The expected-work leader order is a permutation of the residue groups; plans stay valid
    (leaders before the atoms that depend on them, anchors earlier than their position)."""
    sample = active_templates[::40]
    from helpers import default_distances
    dist = default_distances(sample)
    new = CompiledLibrary(sample, 2.0, dist, dist)
    old = CompiledLibrary(sample, 2.0, dist, dist, plan_order="leaders_first_greedy")
    changed = 0
    for lib in (new, old):
        for t in range(len(sample)):
            a0, a1 = int(lib.atom_off[t]), int(lib.atom_off[t + 1])
            order = lib.plan_atom[a0:a1].tolist()
            assert sorted(order) == list(range(a1 - a0))
            src, anchor = lib.plan_src[a0:a1], lib.plan_anchor[a0:a1]
            assert src[0] < 0
            n_leaders = int((src < 0).sum())
            assert all(s < 0 for s in src[:n_leaders]) and all(0 <= s < n_leaders for s in src[n_leaders:])
            assert all(int(anchor[k]) < k for k in range(1, a1 - a0))
    for t in range(len(sample)):
        a0, a1 = int(new.atom_off[t]), int(new.atom_off[t + 1])
        changed += not np.array_equal(new.plan_atom[a0:a1], old.plan_atom[a0:a1])
        # same leaders, possibly another order
        n_leaders = int((new.plan_src[a0:a1] < 0).sum())
        assert sorted(new.plan_atom[a0:a0 + n_leaders].tolist()) == sorted(old.plan_atom[a0:a0 + n_leaders].tolist())
    assert changed > 0
    with pytest.raises(ValueError):
        CompiledLibrary(sample[:2], 2.0, 1.5, 1.5, plan_order="alphabetical")


def _oracle_records(matcher, molecules):
    """Hit records as the device would return them, computed by the CPU oracle (test infrastructure):
    lets the Matcher's host logic -- size groups, completeness, filtering, ordering -- run without a GPU."""
    import oracle
    from enzymm_b200.engine import HIT_DTYPE, HIT_NO_MODEL, HIT_ORIENTED, HIT_PASS
    matcher._compile()
    rows = []
    for size, lo, hi in matcher._groups:
        group = matcher._ordered[lo:hi]
        rmsd, dist, dyn = matcher._get_jess_parameters(size)
        per_mol = oracle.query(molecules, oracle.OracleTemplates(group), rmsd, dist, dyn, max_candidates=10000,
                               ignore_chain=True, threads=4)
        for mi, hits in enumerate(per_mol):
            for h in hits:
                t = group[h.template_index]
                rec = np.zeros((), dtype=HIT_DTYPE)
                rec["structure"], rec["template_index"] = mi, lo + h.template_index
                rec["n_complete"], rec["n_atoms"] = h.n_complete, len(h.atoms)
                rec["rmsd"], rec["rot"], rec["qbar"], rec["tbar"] = h.rmsd, h.rot.reshape(9), h.qbar, h.tbar
                rec["atoms"][:len(h.atoms)] = h.atoms
                flags = 0
                if getattr(t, "residues", None):
                    orient = oracle.orientation(t, h.transform(molecules[mi].xyz[h.atoms]))
                    rec["orientation"] = orient
                    flags |= HIT_ORIENTED
                    try:
                        if oracle.predicted_correct(t.effective_size, dist, h.rmsd, orient):
                            flags |= HIT_PASS
                    except KeyError:
                        flags |= HIT_NO_MODEL
                rec["flags"] = flags
                rows.append(rec)
    records = np.array(rows, dtype=HIT_DTYPE) if rows else np.zeros(0, dtype=HIT_DTYPE)
    return records[np.lexsort((records["template_index"], records["structure"]))]


def test_matcher_host_logic_on_oracle_records(mol_1amy):
    """Reference ``TestMatcher`` expectations through ``Matcher._assemble`` with oracle-computed
    records: grouping by size, completeness before filtering, filter verdicts, result order."""
    from enzymm_b200 import jess_run
    from enzymm_b200.templates import load_templates
    from helpers import oracle_matcher_run
    mol_af = Molecule.load(GOLDEN / "AF-P0DUB6-F1-model_v4.pdb")
    res5 = list(load_templates(subset="5_residues/results/csa3d_0285/"))
    res4 = list(load_templates(subset="4_residues/results/csa3d_0285/"))
    res3 = list(load_templates(subset="3_residues/results/csa3d_0344/"))
    molecules = [mol_1amy, mol_af]
    for kwargs in (dict(), dict(filter_matches=False)):
        matcher = jess_run.Matcher(templates=res5 + res4 + res3, **kwargs)
        got = matcher._assemble(_oracle_records(matcher, molecules), molecules)
        want = oracle_matcher_run(res5 + res4 + res3, molecules, **kwargs)
        assert [molecules.index(k) for k in got] == list(want)
        for mi, expect in want.items():
            mine = got[molecules[mi]]
            assert [m.hit.template.id for m in mine] == [m.template.id for m in expect]
            assert [m.complete for m in mine] == [m.complete for m in expect]
            assert [m.hit.atom_indices for m in mine] == [m.hit.atoms for m in expect]
            for m in mine:
                assert m.predicted_correct == m.hit.device_pass          # reference formula vs record flag
                assert m.orientation == pytest.approx(m.hit.orientation, abs=1e-9)
    m1 = jess_run.Matcher(templates=res5 + res4)
    out = m1._assemble(_oracle_records(m1, molecules), molecules)
    assert len(out[mol_1amy]) == 2 and len(out[mol_af]) == 2           # reference test_matcher.py counts
    assert [m.query_residue_count for m in out[mol_af]] == [511, 511]


def test_pack_files_columns_outlive_the_batch(active_templates):
    """The packed columns are views of native buffers; the buffers live as long as any view does."""
    import gc
    from enzymm_b200.packing import pack_files
    lib = CompiledLibrary(active_templates[:5], 2.0, 1.5, 1.5)
    batch, _ = pack_files([GOLDEN / "1AMY.pdb"] * 4, lib, threads=2)
    xyz, klass = batch.xyz, batch.klass
    want_sum, want_classes = float(xyz.sum()), klass.copy()
    del batch
    gc.collect()
    junk = [np.ones(1 << 20) for _ in range(8)]          # churn the allocator
    assert float(xyz.sum()) == want_sum and np.array_equal(klass, want_classes)
    del junk


def test_rank_order_reference_vectors():
    """``utils.ranked_argsort`` known answers (reference tests/test_utils.py:10-15); feeds
    ``Match.preserved_resid_order``."""
    from enzymm_b200.templates import rank_order
    assert rank_order([0, 4, 8, 6]) == [1, 2, 4, 3]
    assert rank_order([2, 3, 20, 9]) == [1, 2, 4, 3]
    assert rank_order([-3, 3, 20, 9]) == [1, 2, 4, 3]
    assert rank_order([2, 3, 20, 20, 9]) == [1, 2, 4, 4, 3]
    assert rank_order([2, 20, 3, 20, 9]) == [1, 4, 2, 4, 3]


def test_vec3_reference_vectors():
    """Known answers of the reference's ``TestVec3`` (tests/test_template.py:55-167)."""
    vec1, vec2, vec3 = Vec3(1, 2, 3), Vec3(-0.275, 2.8, 0.837), Vec3(0, 0, 0)
    vec4, vec5 = Vec3(1.00000000001, 0, 0), Vec3(-1, 0, 0)
    vec6 = Vec3(-0.43667809853452577, 0.6652199071133453, 0.6056357927338702)
    vec7 = Vec3(-0.4366780985345203, 0.6652199071133459, 0.6056357927338736)
    notavec = (0, 3, 5)
    assert (vec1.x, vec1.y, vec1.z) == (1.0, 2.0, 3.0) and (vec2.x, vec2.y, vec2.z) == (-0.275, 2.8, 0.837)
    assert vec1.norm == math.sqrt(14) and vec3.norm == 0.0
    n = vec1.normalize()
    assert (n.x, n.y, n.z) == (1 / math.sqrt(14), 2 / math.sqrt(14), 3 / math.sqrt(14)) and vec3.normalize() == vec3
    for got, want in (((vec1 + vec2), (0.725, 4.8, 3.837)), ((vec1 + 5), (6, 7, 8)), ((vec1 - vec2), (1.275, -0.8, 2.163)),
                      ((vec1 - 5), (-4, -3, -2)), ((vec1 / vec2), (-3.6363636, 0.7142857, 3.58422939)),
                      ((vec1 / 2), (0.5, 1, 1.5))):
        assert (got.x, got.y, got.z) == pytest.approx(want, abs=5e-8)
    assert vec1 + vec3 == vec1 and vec1 - vec3 == vec1
    with pytest.raises(ZeroDivisionError):
        vec1 / vec3
    for op in (lambda: vec1 + notavec, lambda: vec1 - notavec, lambda: vec1 / notavec, lambda: vec1 @ notavec):
        with pytest.raises(TypeError):
            op()
    assert vec1 @ vec3 == 0 and vec1 @ vec2 == pytest.approx(7.836)
    assert vec1.normalize() @ vec2.normalize() == pytest.approx(0.713465, abs=5e-7)
    assert vec1.angle_to(vec2) == pytest.approx(math.acos(0.713465), abs=5e-7)
    assert vec4.angle_to(vec4) == pytest.approx(0, abs=1e-7) and vec4.angle_to(vec5) == pytest.approx(math.pi, abs=1e-7)
    assert vec6.angle_to(vec7) == pytest.approx(0, abs=1e-7)          # cosine rounds above 1: clamped, not a math error


def test_residue_reference_attributes(all_templates):
    """Reference ``TestResidue.test_attributes`` (tests/test_template.py:709-754): residue typing and
    the orientation vectors the device filter is fed with (GLU: C -> midpoint of the Os; HIS: CG -> ND1)."""
    from enzymm_b200.templates import check_template
    t = next(t for t in all_templates
             if t.pdb_id == "1b74" and len(t.residues) == 6 and t.residues[0].residue_number == 147
             and t.cluster is not None and (t.cluster.id, t.cluster.member, t.cluster.size) == (1, 1, 1))
    r1, r2 = t.residues[0], t.residues[1]
    assert (r1.residue_name, r1.allowed_residues, r1.match_mode, r1.backbone, r1.residue_number, r1.chain_id) == \
           ("GLU", "E", 3, False, 147, "A")
    v = r1.orientation_vector
    assert (v.x, v.y, v.z) == pytest.approx((-0.125, -0.347499999, 0.45599999), abs=5e-8)
    assert r1.orientation_vector_indices == (0, 9)
    v = r2.orientation_vector
    assert (v.x, v.y, v.z) == pytest.approx((1.079, -0.638, 0.57), abs=5e-8)
    assert r2.orientation_vector_indices == (0, 1)
    assert check_template(t, warn=False) is True and check_template(t, warn=True) is True


def test_template_reference_good_loads():
    """Reference ``TestTemplate.test_good_loads`` / ``test_copy`` / ``test_template_non_equality``
    (tests/test_template.py:335-440) on the same two shipped templates."""
    from enzymm_b200.templates import iter_bundle
    texts = {name: text for name, text in iter_bundle() if "1b74_A147-AA180-AA70-AA178-AA8-AA7" in name
             or "csa3d_0011.cluster_1_1_3.1qum_D145" in name}
    text1 = next(v for k, v in texts.items() if "6_residues/results/csa3d_0001/csa3d_0001.cluster_1_1_1.1b74" in k)
    text2 = next(v for k, v in texts.items() if "4_residues/results/csa3d_0011/csa3d_0011.cluster_1_1_3.1qum" in k)
    t1 = Template.loads(text1, warn=True)
    t1_with_id = Template.loads(text1, id="hello_world", warn=True)
    t2 = Template.loads(text2, id="hello_world", warn=True)
    want1 = dict(pdb_id="1b74", template_id_string="1b74_A147-AA180-AA70-AA178-AA8-AA7", mcsa_id=1, uniprot_id="P56868",
                 organism="Aquifex pyrophilus", organism_id="2714", resolution=2.3, experimental_method="X-ray diffraction",
                 ec=("5.1.1.3",), represented_sites=2, enzyme_discription="GLUTAMATE RACEMASE (E.C.5.1.1.3)",
                 effective_size=6, dimension=6, multimeric=True, relative_order=[0], cath=("3.40.50.1860",))
    want2 = dict(pdb_id="1qum", id="hello_world",
                 template_id_string="1qum_D145-D109-D37-D72-D69-D229-D182-D231-D261-D216-D179", mcsa_id=11, uniprot_id=None,
                 organism="Escherichia coli", organism_id="562", resolution=1.55, experimental_method="X-ray diffraction",
                 ec=("3.1.21.2",), represented_sites=1, enzyme_discription="ENDONUCLEASE IV (E.C.3.1.21.2)/DNA",
                 dimension=4, effective_size=4, multimeric=False, relative_order=[2, 1, 4, 3], cath=("3.20.20.150",))
    for t, want in ((t1, want1), (t2, want2)):
        for key, value in want.items():
            assert getattr(t, key) == value, (key, getattr(t, key), value)
    assert (t1.cluster.id, t1.cluster.member, t1.cluster.size) == (1, 1, 1) and len(t1.residues) == 6
    assert (t2.cluster.id, t2.cluster.member, t2.cluster.size) == (1, 1, 3) and len(t2.residues) == 4
    c = t1.copy()
    assert c == t1 and list(c) == list(t1) and hash(c) == hash(t1) and c.cluster == t1.cluster
    assert t1 != t1_with_id and t1 != t2


def test_cluster_reference():
    """Reference ``TestCluster`` (tests/test_template.py:165-178)."""
    from enzymm_b200.templates import Cluster
    with pytest.raises(ValueError):
        Cluster(1, 2, 1)                       # member index beyond the cluster size
    c = Cluster(3, 1, 2)
    assert (c.id, c.member, c.size) == (3, 1, 2)


def _oracle_hit_record(template, molecule, rmsd, dist, dyn):
    """One device-shaped hit record computed by the oracle (test infrastructure only)."""
    import oracle
    from enzymm_b200.engine import HIT_DTYPE, HIT_NO_MODEL, HIT_ORIENTED, HIT_PASS
    (h,) = oracle.query([molecule], oracle.OracleTemplates([template]), rmsd, dist, dyn, max_candidates=10000,
                        ignore_chain=True)[0]
    rec = np.zeros((), dtype=HIT_DTYPE)
    rec["n_complete"], rec["n_atoms"] = h.n_complete, len(h.atoms)
    rec["rmsd"], rec["rot"], rec["qbar"], rec["tbar"] = h.rmsd, h.rot.reshape(9), h.qbar, h.tbar
    rec["atoms"][:len(h.atoms)] = h.atoms
    rec["orientation"] = oracle.orientation(template, h.transform(molecule.xyz[h.atoms]))
    try:
        verdict = HIT_PASS if oracle.predicted_correct(template.effective_size, dist, h.rmsd, float(rec["orientation"])) else 0
    except KeyError:                 # no logistic models for this distance (jess_run.py:339-342)
        verdict = HIT_NO_MODEL
    rec["flags"] = HIT_ORIENTED | verdict
    return rec


def test_match_and_writers_on_oracle_records(mol_1amy):
    """The reference's ``TestMatch`` known answers and golden files (tests/test_jess_run.py:61-179)
    through the product's ``Hit`` / ``Match`` host code, fed with oracle-computed records (no GPU)."""
    import io
    from enzymm_b200 import jess_run
    from enzymm_b200.pyjess_api import Hit
    from test_oracle_golden import T1_PATH, T2_PATH, bundle_templates
    t1, t2 = bundle_templates([T1_PATH, T2_PATH])
    match1 = jess_run.Match(hit=Hit(_oracle_hit_record(t1, mol_1amy, 2, 1.5, 1.5), t1, mol_1amy), pairwise_distance=1.5,
                            complete=True, index=0)
    assert match1.hit.molecule().id == "1AMY" and match1.index == 0
    assert match1.query_atom_count == 3339 and match1.query_residue_count == 403
    assert match1.hit.rmsd == pytest.approx(0.32093143, abs=5e-8)
    assert match1.orientation == pytest.approx(0.15327054322, abs=5e-8)
    expected = [(0.2290067979141952, -0.3853409610281773, 0.377114677867322),
                (0.4249816660862038, -0.21966898402981627, -0.3540863184957992),
                (0.45459385444007694, -0.34869961601989985, 0.10687378206512577),
                (-0.8733960645698886, 0.2563504028143271, -0.9840695023070225),
                (-0.510183600042339, -0.1958417994791759, 0.18963368325429997)]
    assert [(v.x, v.y, v.z) for v in match1.match_vector_list] == [pytest.approx(e, rel=1e-9, abs=1e-9) for e in expected]
    assert match1.template_vector_list == [r.orientation_vector for r in t1.residues]
    assert match1.preserved_resid_order is True and match1.multimeric is False and match1.complete is True
    assert match1.matched_residues == [("GLU", "A", "204"), ("ASP", "A", "87"), ("ASP", "A", "179"),
                                       ("HIS", "A", "288"), ("ASP", "A", "289")]
    assert match1.predicted_correct is True and match1.hit.device_pass
    match2 = jess_run.Match(hit=Hit(_oracle_hit_record(t2, mol_1amy, 2, 1, 1), t2, mol_1amy))
    assert match2.hit.rmsd == pytest.approx(1.7353479120, abs=5e-8)
    assert match2.orientation == pytest.approx(1.6503123465442575, abs=1e-9)
    assert match2.preserved_resid_order is False and match2.complete is False and match2.multimeric is False
    assert match2.matched_residues == [("TRP", "A", "38"), ("HIS", "A", "288"), ("ASP", "A", "289")]
    # writers: byte parity with the reference's golden files
    for name, kwargs in (("1AMY_matches_no_query.pdb", dict(transform=False, include_query=False)),
                         ("1AMY_matches_query_included.pdb", dict(transform=False, include_query=True)),
                         ("1AMY_matches_template.pdb", dict(transform=True, include_query=False))):
        buffer = io.StringIO()
        match1.dump2pdb(buffer, **kwargs)
        assert buffer.getvalue() == (GOLDEN / name).read_text(), name
    got_header, got_row = [l.split("\t") for l in match1.dumps(header=True).splitlines()]
    want_header, want_row = [l.split("\t") for l in (GOLDEN / "results.tsv").read_text().splitlines()]
    assert got_header == want_header
    skip = {"log_evalue", "number_of_mutated_residues", "number_of_side_chain_residues_(template,reference)",
            "number_of_metal_ligands_(template,reference)", "number_of_ptm_residues_(template, reference)",
            "total_reference_residues"}
    assert [(c, g) for c, g, w in zip(want_header, got_row, want_row) if c not in skip and g != w] == []


def test_remark_annotation_parsing_reference():
    """Reference ``test_annotation_parsing`` (tests/test_template.py:509-541) on the REMARK handlers."""
    from enzymm_b200 import templates as T
    h = T._REMARK_HANDLERS
    for tag, token in (("PDB_ID", "abcde"), ("UNIPROT_ID", "abcde"), ("MCSA_ID", "abcde"), ("CLUSTER", "abcde"),
                       ("CLUSTER", "ab_cd_e"), ("RESOLUTION", "abcde"), ("REPRESENTING", "abcde")):
        with pytest.raises(ValueError):
            h[tag]([0, 1, token], {}, True)
    for token in ("1.1.1", "8.1.1.1"):
        with pytest.raises(ValueError):
            h["EC"]([0, 1, token], {"ec": []}, True)
    with pytest.warns(Warning):
        meta = {"ec": []}
        h["EC"]([0, 1, "1.1.1.n1"], meta, True)
    assert meta["ec"] == ["1.1.1.n1"]
    with pytest.raises(ValueError):
        h["CATH"]([0, 1, "1.1.1.n1"], {"cath": []}, True)
    for token in ("1.1.1.1", "1.1.1800.1"):
        meta = {"cath": []}
        h["CATH"]([0, 1, token], meta, True)
        assert meta["cath"] == [token]


def test_matcher_init_reference():
    """Reference ``TestMatcher.test_init`` (tests/test_jess_run.py:278-299): defaults, size detection,
    cpu counts, duplicate detection -- none of it needs a device."""
    import os
    from enzymm_b200 import jess_run
    from enzymm_b200.templates import load_templates
    res5 = list(load_templates(subset="5_residues/results/csa3d_0285/"))
    res4 = list(load_templates(subset="4_residues/results/csa3d_0285/"))
    res3 = list(load_templates(subset="3_residues/results/csa3d_0344/"))
    defaults = {3: {"rmsd": 2, "distance": 0.9, "max_dynamic_distance": 0.9},
                4: {"rmsd": 2, "distance": 1.7, "max_dynamic_distance": 1.7},
                5: {"rmsd": 2, "distance": 2.0, "max_dynamic_distance": 2.0},
                6: {"rmsd": 2, "distance": 2.0, "max_dynamic_distance": 2.0},
                7: {"rmsd": 2, "distance": 2.0, "max_dynamic_distance": 2.0},
                8: {"rmsd": 2, "distance": 2.0, "max_dynamic_distance": 2.0}}
    m1 = jess_run.Matcher(templates=res5 + res4, cpus=2, warn=True)
    m2 = jess_run.Matcher(templates=res5 + res4, skip_smaller_hits=True)
    with pytest.warns(Warning):
        m3 = jess_run.Matcher(templates=res5 + res4 + res3, match_small_templates=True, warn=True, cpus=-1)
    available = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    assert m1.jess_params == defaults and m1.template_effective_sizes == [5, 4]
    assert m1.cpus == 2 and m2.cpus == available and m3.cpus == max(1, available - 1)
    with pytest.raises(ValueError):
        jess_run.Matcher(templates=res5 + res5)
    assert m1.run([]) == {} and m1._engine is None          # nothing to search: no device touched


def _random_pdb_text(rng) -> str:
    """A syntactically valid but unruly PDB file: odd chain ids, negative residue numbers, split
    residues, HETATM records, blank or missing B-factors, short lines, CRLF line ends, a second model."""
    names = ["N", "CA", "C", "O", "CB", "CG", "OD1", "ND2", "OG", "SG", "NE2", "ZN", "OXT", "H1"]
    residues = ["ALA", "ASP", "HIS", "SER", "CYS", "GLU", "HOH", "MSE", "ZN", "UNK", "A"]
    chains = ["A", "B", " ", "AA", "1", "z"]
    lines, serial = ["HEADER    TEST" + " " * 48 + "1ABC"], 1
    blocks = []
    for _ in range(int(rng.integers(1, 40))):
        chain, resname, resnum = chains[rng.integers(len(chains))], residues[rng.integers(len(residues))], int(rng.integers(-20, 400))
        block = []
        for name in rng.choice(names, size=int(rng.integers(1, 8)), replace=False):
            x, y, z = rng.uniform(-99, 999, 3)
            rec = "HETATM" if resname in ("HOH", "ZN", "MSE") and rng.random() < 0.7 else "ATOM  "
            core = f"{rec}{serial % 100000:5d} {name:<4s}{' ' if rng.random() < 0.9 else 'B'}{resname:>3s}{chain:>2s}{resnum:4d}" \
                   f"{' ' if rng.random() < 0.95 else 'A'}   {x:8.3f}{y:8.3f}{z:8.3f}"
            tail = rng.integers(4)
            if tail == 0:
                line = core                                             # stops after z
            elif tail == 1:
                line = core + f"{rng.uniform(0, 1):6.2f}"              # occupancy only
            elif tail == 2:
                line = core + f"{rng.uniform(0, 1):6.2f}{rng.uniform(0, 99):6.2f}"
            else:
                line = (core + f"{1.0:6.2f}{rng.uniform(0, 99):6.2f}").ljust(76) + f"{name[0]:>2s}" + ("1-" if rng.random() < 0.1 else "  ")
            block.append(line)
            serial += 1
        blocks.append(block)
    if len(blocks) > 3 and rng.random() < 0.5:                          # split one residue around another
        first = blocks[0]
        if len(first) > 1:
            blocks[0], extra = first[:1], first[1:]
            blocks.insert(2, extra)
    for block in blocks:
        lines.extend(block)
        if rng.random() < 0.1:
            lines.append("TER")
    if rng.random() < 0.3:
        lines += ["ENDMDL", "MODEL        2", "ATOM      1  N   ALA A   1       0.000   0.000   0.000  1.00  0.00           N", "ENDMDL"]
    lines.append("END")
    return ("\r\n" if rng.random() < 0.2 else "\n").join(lines) + "\n"


def test_ingest_paths_agree_on_unruly_files(tmp_path, active_templates):
    """Three readers, one answer: native packer == native Molecule reader + ``pack_molecules`` ==
    pure-Python parser + ``pack_molecules``, on randomly generated awkward-but-valid files."""
    from enzymm_b200.packing import pack_files
    from enzymm_b200.structures import _parse_pdb_text, load_many
    lib = CompiledLibrary(active_templates[::60], 2.0, 1.5, 1.5)
    rng = np.random.default_rng(20230210)
    paths = []
    for i in range(40):
        p = tmp_path / f"unruly_{i}.pdb"
        p.write_bytes(_random_pdb_text(rng).encode())
        paths.append(p)
    native, ids = pack_files(paths, lib, threads=3)
    through_molecules = pack_molecules(load_many(paths, threads=2), lib)
    python_mols = []
    for p in paths:
        cols, xyz, hid = _parse_pdb_text(p.read_bytes().decode().splitlines(True))
        python_mols.append(Molecule._from_columns(cols, xyz, hid))
    through_python = pack_molecules(python_mols, lib)
    _assert_batches_equal(native, through_molecules)
    _assert_batches_equal(native, through_python)
    assert ids == [m.id for m in python_mols] == ["1ABC"] * len(paths)
    assert [m == n for m, n in zip(load_many(paths), python_mols)] == [True] * len(paths)


def test_native_and_numpy_pack_molecules_agree(tmp_path, active_templates, mol_1amy):
    """``pack_molecules`` takes the native route for molecules that still hold raw name bytes and the
    NumPy route otherwise: same batch, including masked molecules, an empty one and a split residue."""
    from enzymm_b200.structures import _COLUMNS, load_many
    lib = CompiledLibrary(active_templates[::50], 2.0, 1.5, 1.5)
    lines = [l for l in (GOLDEN / "1AMY.pdb").read_text().splitlines() if l.startswith(("ATOM", "HETATM"))]
    first = [l for l in lines if int(l[22:26]) == 1]
    split = tmp_path / "split.pdb"
    split.write_text("\n".join(first[:3] + [l for l in lines if 1 < int(l[22:26]) <= 3] + first[3:] + lines[40:120]) + "\n")
    empty = tmp_path / "empty.pdb"
    empty.write_text("END\n")
    native_mols = load_many([GOLDEN / "1AMY.pdb", split, empty, GOLDEN / "AF-P0DUB6-F1-model_v4.pdb"])
    native_mols.append(native_mols[3].conserved(80))
    native_mols.append(native_mols[0][10:500])
    assert all(m._cols.raw for m in native_mols)
    plain = [Molecule._from_columns({k: m.column(k) for k, _ in _COLUMNS}, m.xyz, m.id) for m in native_mols]
    assert not any(m._cols.raw for m in plain) and plain == native_mols
    for with_chain in (True, False):
        _assert_batches_equal(pack_molecules(native_mols, lib, with_chain=with_chain),
                              pack_molecules(plain, lib, with_chain=with_chain))
    mixed = [native_mols[0], plain[1], native_mols[3]]           # one plain molecule: the NumPy route for all
    _assert_batches_equal(pack_molecules(mixed, lib), pack_molecules([plain[0], plain[1], plain[3]], lib))


def test_classify_keys_including_slot_collisions(active_templates):
    """The vectorised kind -> class lookup equals ``class_of`` kind by kind, also when the hash table
    is so small that most kinds lose their slot and are answered from the overflow map."""
    from enzymm_b200.chem import RESIDUE_ATOMS
    kinds = [(res, name) for res, names in RESIDUE_ATOMS.items() for name in names] + [("HOH", "O"), ("ZN", "ZN"), ("XYZ", "Q1")]

    def key_of(res, name):
        pack = lambda text: int.from_bytes(text.encode().ljust(4, bytes(1)), "little")
        return (pack(res) << 32) | pack(name)

    rng = np.random.default_rng(3)
    picks = rng.integers(0, len(kinds), 5000)
    keys = np.asarray([key_of(*kinds[i]) for i in picks], dtype=np.uint64)
    for bits in (16, 3):
        lib = CompiledLibrary(active_templates[::80], 2.0, 1.5, 1.5)
        lib._HASH_BITS = bits
        want = np.asarray([lib.class_of(*kinds[i]) for i in picks], dtype=np.uint16)
        for _ in range(2):                       # second call: everything answered from the table / overflow map
            assert np.array_equal(lib.classify_keys(keys), want)
        if bits == 3:
            assert len(lib._kind_table[2]) > 0   # the overflow map really was used


def test_residue_ordinals_against_a_plain_python_reference():
    """Residues are numbered by first appearance of (chain, residue number); split residues are
    regrouped by a stable sort.  Checked against a dictionary-based restatement on random inputs."""
    rng = np.random.default_rng(11)
    for trial in range(200):
        n = int(rng.integers(1, 60))
        chain = rng.integers(0, 3, n).astype(np.uint16)
        resnum = rng.integers(-3, 6, n).astype(np.int32)
        if trial % 2:                                   # mostly contiguous residues, as in real files
            order = np.lexsort((resnum, chain))
            chain, resnum = chain[order], resnum[order]
        first_seen = {}
        want = []
        for c, r in zip(chain.tolist(), resnum.tolist()):
            want.append(first_seen.setdefault((c, r), len(first_seen)))
        contiguous = all(want[i] >= want[i - 1] for i in range(1, n))
        ordinal, order = residue_ordinals(chain, resnum)
        if contiguous:
            assert order is None and ordinal.tolist() == want
        else:
            assert order is not None and sorted(order.tolist()) == list(range(n))
            assert ordinal.tolist() == sorted(want)                         # grouped, first-appearance numbering
            assert [want[i] for i in order.tolist()] == ordinal.tolist()    # the permutation realises it
            for a, b in zip(order.tolist(), order.tolist()[1:]):            # stable inside a residue
                assert want[a] != want[b] or a < b


def test_multi_device_scan_hands_chunks_back_in_input_order(monkeypatch, active_templates):
    """``Matcher.scan_files(devices=[...])``: one worker thread per device pulls chunks from one counter;
    the generator returns them in input order whatever order they finish in, propagates a worker's
    exception, and lets the workers stop when it is abandoned early.  (Device work replaced by a stub:
    this is the host-side hand-out and re-ordering; the GPU test runs the real thing.)"""
    import random
    import threading
    import time
    from enzymm_b200 import jess_run

    matcher = jess_run.Matcher(active_templates[:50])
    paths = [f"/nowhere/s{i:04d}.pdb" for i in range(103)]
    seen_threads = set()
    real = jess_run.Matcher.scan_files

    def stub(self, paths, chunk_size=2048, threads=0, queue=None, with_batch=False, devices=None, _with_span=False,
             on_error="raise"):
        if devices is not None:
            yield from real(self, paths, chunk_size, threads, queue, with_batch, devices, _with_span)
            return
        seen_threads.add(threading.get_ident())
        for span in queue:
            time.sleep(random.random() * 0.01)
            if getattr(self, "_explode_at", None) == span[0]:
                raise RuntimeError("device lost")
            yield (paths[span[0]:span[1]], [None] * (span[1] - span[0]), np.zeros(span[0], dtype=np.int8), span)

    monkeypatch.setattr(jess_run.Matcher, "scan_files", stub)
    monkeypatch.setattr(jess_run.Matcher, "close", lambda self: None)
    chunks = list(matcher.scan_files(paths, chunk_size=10, devices=[0, 1, 2]))
    assert [c[0] for c in chunks] == [paths[i:i + 10] for i in range(0, 103, 10)]
    assert [len(c[2]) for c in chunks] == list(range(0, 103, 10)) and all(len(c) == 3 for c in chunks)
    assert len(seen_threads) == 3
    # abandoned after two chunks: the generator's close must not hang on workers blocked in the queue
    gen = matcher.scan_files(paths * 20, chunk_size=5, devices=[0, 1])
    assert [next(gen)[0][0], next(gen)[0][0]] == [paths[0], paths[5]]
    t0 = time.time()
    gen.close()
    assert time.time() - t0 < 5
    # a failing worker surfaces in the consumer
    matcher._device_workers = {}
    monkeypatch.setattr(jess_run.Matcher, "_explode_at", 40, raising=False)
    with pytest.raises(RuntimeError, match="device lost"):
        list(matcher.scan_files(paths, chunk_size=10, devices=[0, 1]))
