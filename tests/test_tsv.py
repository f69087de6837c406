"""The native results-table writer (``tsv.TableWriter`` + ``emm_tsv_format``) against the reference's
own row writer semantics: ``Match.dump`` (``enzymm/jess_run.py:185-284``) driven as ``_cli.py:270-316``
drives it.  Hit records come from the CPU oracle here (host logic; no GPU needed)."""
import ctypes
import io
import math
import random

import numpy as np
import pytest

from conftest import GOLDEN
from enzymm_b200 import jess_run
from enzymm_b200.engine import load_cdll
from enzymm_b200.packing import pack_files, pack_molecules
from enzymm_b200.synth import SynthConfig, generate_chunk
from enzymm_b200.tsv import TableWriter
from test_host_model import _oracle_records


def test_float_columns_print_as_python_prints_them():
    """``str(round(x, 5))`` -- shortest round-trip repr of the correctly rounded value -- for the
    values a table can hold and the formats repr switches between."""
    lib = load_cdll()
    lib.emm_tsv_repr_round5.restype = ctypes.c_int
    buf = ctypes.create_string_buffer(64)
    rng = random.Random(7)
    values = [0.0, -0.0, 1.0, 2.0, 0.5, 0.1, 1e-5, 4.9e-6, 5e-6, 1.5e-5, 2.675, 0.000125, 0.32093143, 1.7353479120,
              0.15327054322, 1.6503123465442575, math.pi, 123456.789012, 1e16, 1.23456789e17, -3.08424478, 1e-4,
              0.99999499999, 0.999995, 0.9999949999999999, 2.5e-5, 3.5e-5, float("nan"), float("inf"), -float("inf")]
    values += [rng.uniform(0, 3.2) for _ in range(3000)] + [rng.uniform(0, 1e-3) for _ in range(500)]
    values += [round(rng.uniform(0, 3), 5) + 5e-6 for _ in range(500)] + [10 ** rng.uniform(-8, 18) for _ in range(500)]
    for v in values:
        assert lib.emm_tsv_repr_round5(ctypes.c_double(v), buf, 64) == 0
        assert buf.value.decode() == str(round(v, 5)), v


def _reference_rows(matcher, records, molecules) -> str:
    """The table as the reference CLI writes it (``_cli.py:270-316``): Match objects, one dump each."""
    out = io.StringIO()
    for index, (_, matches) in enumerate(matcher._assemble(records, molecules).items()):
        for jndex, match in enumerate(matches):
            match.index = jndex + 1
            match.dump(out, header=(index + jndex == 0))
    return out.getvalue()


def test_native_rows_equal_match_dump(tmp_path, active_templates, mol_1amy, mol_af):
    """Byte for byte: filtered and unfiltered, skip-irrelevant (records decide), real fixtures with
    annotated cluster templates (completeness both ways) + synthetic structures, through both native
    packers (Molecule objects and files)."""
    chunk = generate_chunk(3, SynthConfig(), active_templates, 10)
    paths = []
    for i in range(chunk.n_structures):
        path = tmp_path / f"query_{i % 7}.pdb" if i < 7 else tmp_path / f"dir{i}" / f"query_{i % 7}.pdb"
        path.parent.mkdir(exist_ok=True)
        path.write_text(chunk.to_pdb(i))
        paths.append(path)
    paths += [GOLDEN / "1AMY.pdb", GOLDEN / "AF-P0DUB6-F1-model_v4.pdb"]
    molecules = jess_run.load_molecules(paths)
    assert [m.id for m in molecules][6:9] == ["query_6", "query_0_2", "query_1_2"]
    for kwargs in (dict(), dict(filter_matches=False)):
        matcher = jess_run.Matcher(active_templates, **kwargs)
        records = _oracle_records(matcher, molecules)
        want = _reference_rows(matcher, records, molecules)
        assert len(want.splitlines()) > 20
        writer = TableWriter(matcher)
        from_molecules = pack_molecules(molecules, matcher._compile()).table
        from_files = pack_files(paths, matcher._compile())[0].table
        for table in (from_molecules, from_files):
            got = writer.header() + writer.format(records, table, [m.id for m in molecules]).decode()
            assert got == want
        assert {"True", "False"} <= {line.split("\t")[22] for line in want.splitlines()[1:]}       # completeness both ways
    # predict_correctness=False leaves the column empty (Match.dump(predict_correctness=False))
    matcher = jess_run.Matcher(active_templates, filter_matches=False)
    records = _oracle_records(matcher, molecules[-2:])
    rows = TableWriter(matcher, predict_correctness=False).format(
        records, pack_molecules(molecules[-2:], matcher._compile()).table, ["a\tb", 'q"uote']).decode().splitlines()
    assert all(r.split("\t")[-8] == "" for r in rows) if rows else True
    assert rows[0].startswith('"a\tb"\t') and any(r.startswith('"q""uote"\t') for r in rows)  # csv QUOTE_MINIMAL


def test_reference_golden_row(mol_1amy):
    """``tests/test_data/results.tsv`` of the reference (``test_jess_run.py:147-160``): every column the
    offline build can reproduce -- all but ``log_evalue`` (formula inside the un-vendored Jess, SURVEY
    8c) and the five M-CSA annotation columns (their data blob is absent from the reference checkout)."""
    from test_host_model import _oracle_hit_record
    from test_oracle_golden import T1_PATH, bundle_templates
    (t1,) = bundle_templates([T1_PATH])
    params = {s: {"rmsd": 2, "distance": 1.5, "max_dynamic_distance": 1.5} for s in range(3, 9)}
    matcher = jess_run.Matcher([t1], jess_params=params)
    matcher._compile()
    record = np.array([_oracle_hit_record(t1, mol_1amy, 2, 1.5, 1.5)])
    writer = TableWriter(matcher)
    rows, index, complete, predicted = writer.select(record)
    got = writer.format(record, pack_molecules([mol_1amy], matcher._compile()).table, ["1AMY"],
                        (rows, np.zeros_like(index), np.ones_like(complete), predicted)).decode()
    want_header, want_row = [l.split("\t") for l in (GOLDEN / "results.tsv").read_text().splitlines()]
    got_row = got.rstrip("\n").split("\t")
    assert writer.header().rstrip("\n").split("\t") == want_header
    skip = {"log_evalue", "number_of_mutated_residues", "number_of_side_chain_residues_(template,reference)",
            "number_of_metal_ligands_(template,reference)", "number_of_ptm_residues_(template, reference)",
            "total_reference_residues"}
    assert [(c, g, w) for c, g, w in zip(want_header, got_row, want_row) if c not in skip and g != w] == []
    assert got_row[want_header.index("log_evalue")] == "nan"


def test_missing_model_is_a_keyerror(active_templates, mol_1amy):
    """Filtering at a distance without logistic models raises, as ``jess_run.py:339-342`` does."""
    subset = [t for t in active_templates if t.effective_size == 3][:300]
    params = {s: {"rmsd": 2, "distance": 0.5, "max_dynamic_distance": 0.5} for s in range(3, 9)}
    loose = {s: {"rmsd": 2, "distance": 2.5, "max_dynamic_distance": 2.5} for s in range(3, 9)}
    for p in (params, loose):
        matcher = jess_run.Matcher(subset, jess_params=p)
        records = _oracle_records(matcher, [mol_1amy])
        if len(records):
            with pytest.raises(KeyError):
                TableWriter(matcher).select(records)
            return
    pytest.fail("no hit at a distance without models")
