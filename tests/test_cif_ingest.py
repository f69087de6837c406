"""mmCIF and gzip input of the native ingest (``csrc/emm_pdb.cpp``): the readers tell PDB from mmCIF by
content, as ``pyjess.Molecule.load(format="detect")`` does at the reference's one call site
(``enzymm/jess_run.py:538``), and the path-taking readers inflate gzip-compressed files.

The reference holds no mmCIF or gzip fixture, so nothing here is a reference golden ("unpinned"): the
checks are that an mmCIF rendering of a structure yields exactly the columns its PDB text yields --
through ``Molecule.load``, ``load_many`` and ``pack_files`` -- plus the syntax the CIF grammar allows
around ``_atom_site`` and the documented choices (first model only, ``label_*`` identifiers unless
``use_author``, two-character chain ids)."""
import gzip

import numpy as np
import pytest

from conftest import GOLDEN
from enzymm_b200.library import CompiledLibrary
from enzymm_b200.packing import pack_files, pack_molecules
from enzymm_b200.structures import Molecule, load_many

COLUMNS = ("serial", "name", "altloc", "residue_name", "chain_id", "residue_number", "insertion_code",
           "occupancy", "temperature_factor", "element", "charge")


def cif_value(text: str) -> str:
    text = str(text)
    if text == "":
        return "."
    if any(ch in text for ch in " '\"#_;") or text[0] in "$[]":
        return f'"{text}"' if '"' not in text else f"'{text}'"
    return text


def to_cif(mol: Molecule, block: str = "TEST", models=(1,), with_auth=True, decimals=3, lines=None) -> str:
    """An mmCIF rendering of ``mol``: label_* = the PDB identifiers; auth_* deliberately different
    (chain ``Z``, residue number + 100) so the two readings can be told apart."""
    out = [f"data_{block}", "#", f"_entry.id   {block}", "#",
           "_struct.title", ";A title that spans", "lines; with a ; inside and a loop_ word", ";", "#",
           "loop_", "_citation.id", "_citation.title", "primary 'It''s a title with _tags and loop_'", "#",
           "loop_"]
    tags = ["group_PDB", "id", "type_symbol", "label_atom_id", "label_alt_id", "label_comp_id", "label_asym_id",
            "label_entity_id", "label_seq_id", "pdbx_PDB_ins_code", "Cartn_x", "Cartn_y", "Cartn_z", "occupancy",
            "B_iso_or_equiv", "pdbx_formal_charge"]
    if with_auth:
        tags += ["auth_seq_id", "auth_comp_id", "auth_asym_id", "auth_atom_id"]
    tags.append("pdbx_PDB_model_num")
    out += [f"_atom_site.{t}" for t in tags]
    col = {k: mol.column(k) for k in COLUMNS}
    for model in models:
        for i in range(len(mol)):
            x, y, z = (f"{v:.{decimals}f}" for v in mol.xyz[i])
            row = ["ATOM", col["serial"][i], cif_value(col["element"][i]) if col["element"][i] else "?",
                   cif_value(col["name"][i]), cif_value(col["altloc"][i].strip()), cif_value(col["residue_name"][i]),
                   cif_value(col["chain_id"][i]), "1", col["residue_number"][i],
                   cif_value(col["insertion_code"][i].strip()) if col["insertion_code"][i].strip() else "?",
                   x, y, z, f"{col['occupancy'][i]:.2f}", f"{col['temperature_factor'][i]:.2f}",
                   col["charge"][i] if col["charge"][i] else "?"]
            if with_auth:
                row += [col["residue_number"][i] + 100, cif_value(col["residue_name"][i]), "Z", cif_value(col["name"][i])]
            row.append(model)
            out.append(" ".join(str(v) for v in row))
    out += ["#", "loop_", "_pdbx_poly_seq_scheme.asym_id", "_pdbx_poly_seq_scheme.seq_id", "A 1", "A 2", "#"]
    if lines:
        out += lines
    return "\n".join(out) + "\n"


def assert_same_molecule(got: Molecule, want: Molecule):
    assert len(got) == len(want)
    assert got.xyz.tobytes() == want.xyz.tobytes()
    for key in COLUMNS:
        a, b = got.column(key), want.column(key)
        if a.dtype.kind == "f":
            assert a.tobytes() == b.tobytes(), key
        else:
            assert a.tolist() == b.tolist(), key


@pytest.fixture(scope="module")
def fixture_texts(mol_1amy, mol_af):
    return {"1AMY": (mol_1amy, to_cif(mol_1amy, "1AMY")), "AF": (mol_af, to_cif(mol_af, "AF-P0DUB6-F1"))}


def test_cif_rendering_of_the_fixtures_loads_like_their_pdb_text(fixture_texts):
    for name, (mol, text) in fixture_texts.items():
        got = Molecule.loads(text)
        assert_same_molecule(got, mol)
    assert Molecule.loads(fixture_texts["1AMY"][1]).id == "1AMY"                  # the data block name
    assert Molecule.loads(fixture_texts["AF"][1]).id == "AF-P0DUB6-F1"           # ... in full
    assert Molecule.loads(fixture_texts["AF"][1], id="mine").id == "mine"
    assert Molecule.loads(fixture_texts["1AMY"][1], format="cif").id == "1AMY"
    with pytest.raises(ValueError):
        Molecule.loads(fixture_texts["1AMY"][1], format="pdb")
    with pytest.raises(ValueError):
        Molecule.load(GOLDEN / "1AMY.pdb", format="cif")
    with pytest.raises(ValueError):
        Molecule.load(GOLDEN / "1AMY.pdb", format="xyz")


def test_use_author_reads_the_auth_identifiers(mol_1amy):
    text = to_cif(mol_1amy, "1AMY")
    label, auth = Molecule.loads(text), Molecule.loads(text, use_author=True)
    assert set(label.column("chain_id").tolist()) == set(mol_1amy.column("chain_id").tolist())
    assert set(auth.column("chain_id").tolist()) == {"Z"}
    assert (auth.column("residue_number") == mol_1amy.column("residue_number") + 100).all()
    assert auth.xyz.tobytes() == label.xyz.tobytes()
    # without auth_* items the switch falls back to the label_* ones
    plain = to_cif(mol_1amy, "1AMY", with_auth=False)
    assert_same_molecule(Molecule.loads(plain, use_author=True), mol_1amy)


def test_only_the_first_model_is_read(mol_1amy):
    small = mol_1amy.select(np.arange(len(mol_1amy)) < 50)
    text = to_cif(small, "NMR", models=(1, 2, 3))
    assert_same_molecule(Molecule.loads(text), small)


def test_cif_syntax_around_atom_site(mol_1amy):
    small = mol_1amy.select(np.arange(len(mol_1amy)) < 8)
    # quoted names with an apostrophe, more decimals than a PDB file can hold, comments, CRLF
    text = to_cif(small, "X", decimals=5).replace(" CA ", ' "C1\'" ', 1)
    got = Molecule.loads(text.replace("\n", "\r\n"))
    assert got.column("name")[1] == "C1'" or "C1'" in got.column("name").tolist()
    assert got.xyz.tobytes() == small.xyz.tobytes()               # %.5f of a 3-decimal value: the same doubles
    more = small.with_xyz(small.xyz + 0.00123)
    assert Molecule.loads(to_cif(more, "X", decimals=5)).xyz.tobytes() == np.round(more.xyz, 5).tobytes()
    # item-value pairs instead of a loop: a one-atom category
    pairs = "\n".join(["data_ONE", "_atom_site.group_PDB HETATM", "_atom_site.id 7", "_atom_site.type_symbol ZN",
                       "_atom_site.label_atom_id ZN", "_atom_site.label_comp_id ZN", "_atom_site.label_asym_id B",
                       "_atom_site.label_seq_id .", "_atom_site.auth_seq_id 301", "_atom_site.Cartn_x 1.5",
                       "_atom_site.Cartn_y -2.25", "_atom_site.Cartn_z 1e1", "_atom_site.occupancy 1.00",
                       "_atom_site.B_iso_or_equiv 12.3(4)", ""])
    one = Molecule.loads(pairs)
    assert len(one) == 1 and one.id == "ONE"
    atom = one.atom(0)
    assert (atom.name, atom.residue_name, atom.chain_id, atom.residue_number, atom.serial) == ("ZN", "ZN", "B", 301, 7)
    assert (atom.x, atom.y, atom.z, atom.temperature_factor) == (1.5, -2.25, 10.0, 12.3)
    # a block without atoms is an empty (falsy) molecule, like a PDB file without coordinate records
    assert len(Molecule.loads("data_EMPTY\n_entry.id EMPTY\n")) == 0


def test_cif_errors_are_loud(mol_1amy):
    small = mol_1amy.select(np.arange(len(mol_1amy)) < 4)
    good = to_cif(small, "X")
    with pytest.raises(ValueError, match="chain id"):
        Molecule.loads(good.replace(" A 1 ", " AAA 1 "))
    with pytest.raises(ValueError, match="Cartn"):
        Molecule.loads(good.replace("_atom_site.Cartn_z", "_atom_site.Cartn_w"))
    with pytest.raises(ValueError, match="malformed"):
        Molecule.loads(good.replace(f"{small.xyz[0, 0]:.3f}", "abc", 1))
    with pytest.raises(ValueError, match="middle of a row"):
        Molecule.loads(good[:good.index("#\nloop_\n_pdbx_poly_seq_scheme")].rstrip().rsplit(" ", 1)[0] + "\n")


def test_files_of_both_formats_plain_and_gzipped(tmp_path, fixture_texts, active_templates):
    """``load_many`` and ``pack_files`` over PDB, mmCIF, .pdb.gz and .cif.gz of the same structures."""
    paths, want = [], []
    for name, (mol, cif) in fixture_texts.items():
        pdb = (GOLDEN / ("1AMY.pdb" if name == "1AMY" else "AF-P0DUB6-F1-model_v4.pdb")).read_bytes()
        for suffix, payload in ((".pdb", pdb), (".cif", cif.encode()), (".pdb.gz", gzip.compress(pdb)),
                                (".cif.gz", gzip.compress(cif.encode()[:5000]) + gzip.compress(cif.encode()[5000:]))):
            path = tmp_path / f"{name}{suffix}"
            path.write_bytes(payload)
            paths.append(path)
            want.append(mol)
    for threads in (1, 4):
        mols = load_many(paths, threads=threads)
        assert len(mols) == len(want)
        for got, ref in zip(mols, want):
            assert_same_molecule(got, ref)
    for path, ref in zip(paths, want):
        assert_same_molecule(Molecule.load(path), ref)
    few = active_templates[::40]
    lib = CompiledLibrary(few, 2.0, 1.5, 1.5)
    batch, ids = pack_files(paths, lib, threads=3)
    ref_batch = pack_molecules(want, lib)
    assert batch.atom_off.tolist() == ref_batch.atom_off.tolist()
    assert batch.xyz.tobytes() == ref_batch.xyz.tobytes()
    assert batch.klass.tolist() == ref_batch.klass.tolist()
    assert batch.residue.tolist() == ref_batch.residue.tolist()
    assert batch.bfactor.tobytes() == ref_batch.bfactor.tobytes()
    assert batch.chain.tolist() == ref_batch.chain.tolist()
    assert ids[0] == "1AMY" and ids[1] == "1AMY"                    # HEADER idCode / the data block name
    # the results table's per-structure columns do not depend on the input format either
    assert batch.table.residue_count.tolist()[:4] == [batch.table.residue_count[0]] * 4
    # use_author reaches the packed path: the auth_* rendering has one chain Z, residue numbers + 100
    auth_batch, _ = pack_files(paths[1:2], lib, use_author=True)
    assert set(np.unique(auth_batch.chain).tolist()) == {ord("Z")}


def test_broken_gzip_is_an_error(tmp_path):
    blob = gzip.compress((GOLDEN / "1AMY.pdb").read_bytes())
    path = tmp_path / "cut.pdb.gz"
    path.write_bytes(blob[:len(blob) // 2])
    with pytest.raises(ValueError, match="inflate"):
        load_many([path])
    with pytest.raises(Exception):
        Molecule.load(path)


def test_readers_survive_mutated_inputs_under_sanitizers(tmp_path, mol_1amy):
    """``tests/c/fuzz_readers.cpp`` + ``csrc/emm_pdb.cpp`` under AddressSanitizer / UBSan: 30 000 mutated
    mmCIF and PDB texts and gzip images (truncations, stray quotes / semicolons / keywords, random bytes)
    through ``emm_pdb_count_atoms`` / ``emm_pdb_parse_ex`` and, every eighth one as a file, through
    ``emm_pdb_pack_files_ex`` / ``emm_pdb_load_files_ex`` -- no over-read, no overflow, no crash."""
    import shutil
    import subprocess
    from conftest import ROOT
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    exe = tmp_path / "fuzz_readers"
    build = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                            "-pthread", f"-I{ROOT / 'include'}", "-o", str(exe), str(ROOT / "tests" / "c" / "fuzz_readers.cpp"),
                            str(ROOT / "enzymm_b200" / "csrc" / "emm_pdb.cpp"), "-ldl"],
                           capture_output=True, text=True, timeout=300)
    if build.returncode != 0 and any(word in (build.stderr + build.stdout).lower() for word in ("sanitize", "asan", "ubsan")):
        pytest.skip("this g++ has no sanitizer runtime")
    assert build.returncode == 0, build.stderr
    small = mol_1amy.select(np.arange(len(mol_1amy)) < 60)
    seeds = [tmp_path / "a.cif", tmp_path / "b.cif", tmp_path / "c.pdb"]
    seeds[0].write_text(to_cif(small, "X"))
    seeds[1].write_text(to_cif(small, "NMR", models=(1, 2), with_auth=False, decimals=5))
    seeds[2].write_text("".join((GOLDEN / "1AMY.pdb").read_text().splitlines(keepends=True)[:120]))
    seeds.append(tmp_path / "d.cif.gz")
    seeds[3].write_bytes(gzip.compress(to_cif(small, "Z").encode()))
    run = subprocess.run([str(exe), "30000"] + [str(s) for s in seeds], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and "fuzzed 30000 inputs" in run.stdout, (run.stdout + run.stderr)[-2000:]


def test_bad_files_can_be_skipped_and_named(tmp_path, active_templates, mol_1amy):
    """``pack_files(on_error="skip")``: an unreadable or malformed file does not end a screening run -- it
    is warned about, listed in ``batch.bad_files`` and stays in the batch as a structure without atoms;
    the files around it are packed as if it were not there.  The default still raises, with Python's
    exception types for OS errors (the reference's ``load_molecules`` + ``_cli.py:318-328``)."""
    good = GOLDEN / "1AMY.pdb"
    text = good.read_text()
    broken = tmp_path / "broken.pdb"
    broken.write_text(text.replace(text.splitlines()[700][30:38], "  abcdef", 1))
    cut = tmp_path / "cut.cif.gz"
    cut.write_bytes(gzip.compress(to_cif(mol_1amy, "X").encode())[:2000])
    long_chain = tmp_path / "long_chain.cif"
    small = mol_1amy.select(np.arange(len(mol_1amy)) < 4)
    long_chain.write_text(to_cif(small, "X").replace(" A 1 ", " AAA 1 "))
    folder = tmp_path / "folder.pdb"
    folder.mkdir()
    paths = [good, tmp_path / "missing.pdb", broken, good, cut, long_chain, folder, good]
    lib = CompiledLibrary(active_templates[::40], 2.0, 1.5, 1.5)
    with pytest.raises(FileNotFoundError):
        pack_files(paths, lib)
    with pytest.raises(IsADirectoryError):
        pack_files([good, folder], lib)
    with pytest.raises(ValueError, match="malformed"):
        pack_files([good, broken], lib)
    with pytest.raises(ValueError):
        pack_files(paths, lib, on_error="ignore")
    with pytest.warns(UserWarning) as caught:
        batch, ids = pack_files(paths, lib, threads=3, on_error="skip")
    assert sorted(batch.bad_files) == [1, 2, 4, 5, 6] and len(caught) == 5
    assert "cannot open" in batch.bad_files[1] and "malformed" in batch.bad_files[2] and "inflate" in batch.bad_files[4]
    assert "chain id" in batch.bad_files[5] and "cannot read" in batch.bad_files[6]
    sizes = np.diff(batch.atom_off).tolist()
    assert sizes == [3339, 0, 0, 3339, 0, 0, 0, 3339] and ids == ["1AMY", None, None, "1AMY", None, None, None, "1AMY"]
    alone, _ = pack_files([good], lib)
    for k in (0, 3, 7):
        lo, hi = int(batch.atom_off[k]), int(batch.atom_off[k + 1])
        assert batch.xyz[lo:hi].tobytes() == alone.xyz.tobytes() and batch.klass[lo:hi].tolist() == alone.klass.tolist()
        assert batch.residue[lo:hi].tolist() == alone.residue.tolist()
    assert batch.table.residue_count.tolist() == [alone.table.residue_count[0] if s else 0 for s in sizes]
    clean, _ = pack_files([good, good], lib, on_error="skip")
    assert clean.bad_files == {}
