"""Packed corpus files (``packing.write_corpus`` / ``read_corpus``, ``Matcher.scan_corpus``): structures
parsed and packed once, then screened from the mapped columns -- the "packed binary cache" of
SURVEY.md 8f-2.  Host logic only: the device call is replaced by the CPU oracle where one is needed."""
import io

import numpy as np
import pytest

from conftest import GOLDEN
from enzymm_b200 import jess_run
from enzymm_b200.library import CompiledLibrary
from enzymm_b200.packing import pack_files, read_corpus, slice_batch, write_corpus
from enzymm_b200.synth import SynthConfig, generate_chunk
from test_host_model import _oracle_records
from test_tsv import _reference_rows


def same_batch(got, want):
    assert got.atom_off.tolist() == want.atom_off.tolist()
    assert got.xyz.tobytes() == want.xyz.tobytes()
    for key in ("klass", "residue", "chain"):
        assert getattr(got, key).tolist() == getattr(want, key).tolist(), key
    assert got.bfactor.tobytes() == want.bfactor.tobytes()
    if want.atom_id is not None:
        assert got.atom_id.tolist() == want.atom_id.tolist()
    elif got.atom_id is not None:                 # a slice of a corpus that holds a reordered file elsewhere: identity
        sizes = np.diff(got.atom_off)
        assert got.atom_id.tolist() == (np.arange(got.n_atoms) - np.repeat(got.atom_off[:-1], sizes)).tolist()
    a, b = got.table, want.table
    assert a.kind_names[a.kind].tobytes() == b.kind_names[b.kind].tobytes()        # kind numbering may differ per batch
    assert a.residue.tolist() == b.residue.tolist() and a.res_off.tolist() == b.res_off.tolist()
    assert a.res_key.tolist() == b.res_key.tolist() and a.residue_count.tolist() == b.residue_count.tolist()


@pytest.fixture(scope="module")
def corpus(tmp_path_factory, active_templates):
    root = tmp_path_factory.mktemp("corpus")
    chunk = generate_chunk(5, SynthConfig(), active_templates, 9)
    paths = []
    for i in range(chunk.n_structures):
        path = root / f"q{i % 4}.pdb" if i < 4 else root / f"d{i}" / f"q{i % 4}.pdb"
        path.parent.mkdir(exist_ok=True)
        path.write_text(chunk.to_pdb(i))
        paths.append(path)
    # a file whose residue is split over two runs (atom_id column), and the two fixtures
    lines = (GOLDEN / "1AMY.pdb").read_text().splitlines(keepends=True)
    atoms = [k for k, line in enumerate(lines) if line.startswith("ATOM")]
    moved = lines[:atoms[3]] + lines[atoms[3] + 1:atoms[40]] + [lines[atoms[3]]] + lines[atoms[40]:]
    split = root / "split.pdb"
    split.write_text("".join(moved))
    paths += [split, GOLDEN / "1AMY.pdb", GOLDEN / "AF-P0DUB6-F1-model_v4.pdb"]
    out = root / "all.emmpack"
    assert write_corpus(paths, out, threads=3) == len(paths)
    return paths, out


def test_corpus_round_trip_equals_pack_files(corpus, active_templates):
    paths, out = corpus
    for templates in (active_templates[::40], active_templates[5::97]):           # typing classes are per library
        lib = CompiledLibrary(templates, 2.0, 1.5, 1.5)
        want, want_headers = pack_files(paths, lib)
        got, ids = read_corpus(out, lib)
        same_batch(got, want)
        assert got.atom_id is not None                                            # the split file
        assert got.header_ids == want_headers
        assert ids[:6] == ["q0", "q1", "q2", "q3", "q0_2", "q1_2"] and ids[-3:] == ["split", "1AMY", "AF-P0DUB6-F1-model_v4"]
        for lo, hi in ((0, 4), (3, 10), (9, 12), (11, 12), (5, 5)):
            same_batch(slice_batch(got, lo, hi), pack_files(paths[lo:hi], lib)[0]) if hi > lo else None
            assert slice_batch(got, lo, hi).n_structures == hi - lo
    without_chain, _ = read_corpus(out, lib, with_chain=False)
    assert without_chain.chain is None


def test_corpus_file_errors(tmp_path, corpus, active_templates):
    _, out = corpus
    lib = CompiledLibrary(active_templates[::200], 2.0, 1.5, 1.5)
    with pytest.raises(ValueError, match="not a packed corpus"):
        read_corpus(GOLDEN / "1AMY.pdb", lib)
    cut = tmp_path / "cut.emmpack"
    cut.write_bytes(out.read_bytes()[:100000])
    with pytest.raises(ValueError, match="truncated"):
        read_corpus(cut, lib)
    with pytest.raises(FileNotFoundError):
        write_corpus([tmp_path / "missing.pdb"], tmp_path / "x.emmpack")
    with pytest.raises(ValueError):
        write_corpus([GOLDEN / "1AMY.pdb"], tmp_path / "x.emmpack", ids=["a", "b"])
    empty = tmp_path / "empty.emmpack"
    assert write_corpus([], empty) == 0
    batch, ids = read_corpus(empty, lib)
    assert batch.n_structures == 0 and ids == []
    with pytest.warns(UserWarning):
        assert write_corpus([GOLDEN / "1AMY.pdb", tmp_path / "missing.pdb"], tmp_path / "s.emmpack", on_error="skip") == 2
    batch, ids = read_corpus(tmp_path / "s.emmpack", lib)
    assert np.diff(batch.atom_off).tolist() == [3339, 0] and ids == ["1AMY", "missing"]


def test_scan_corpus_and_table_on_oracle_records(corpus, active_templates, monkeypatch):
    """``scan_files`` / ``scan_to_tsv`` over ``*.emmpack`` paths: chunks of the mapped corpus go to the device
    call (here: the oracle), ids and records come back per chunk, and the table is the one ``Match.dump``
    writes for the same structures, byte for byte."""
    paths, out = corpus
    molecules = jess_run.load_molecules(paths)
    matcher = jess_run.Matcher(active_templates[::3])
    matcher._compile()
    offsets = {}

    def oracle_for(batch):
        # which structures of the corpus is this chunk? (by their first coordinates)
        sizes = np.diff(batch.atom_off).tolist()
        for start in range(len(molecules)):
            window = molecules[start:start + len(sizes)]
            if [len(m) for m in window] == sizes and all(
                    np.array_equal(np.sort(m.xyz[:, 0]), np.sort(batch.xyz[int(batch.atom_off[k]):int(batch.atom_off[k + 1]), 0]))
                    for k, m in enumerate(window)):
                offsets[len(offsets)] = start
                return _oracle_records(matcher, window)
        raise AssertionError("chunk does not match any window of the corpus")

    class FakeSession:                                   # the device side of one lane: upload, run, download
        created = []

        def __init__(self, library, max_atoms, max_structures, hit_capacity):
            self.max_atoms, self.max_structures, self.hit_capacity = max_atoms, max_structures, hit_capacity
            self.batch, self.busy = None, False
            FakeSession.created.append(self)

        def upload(self, batch, stream=0):
            assert not self.busy, "a lane was reused before its chunk was collected"
            assert batch.n_atoms <= self.max_atoms and batch.n_structures <= self.max_structures
            self.batch, self.busy = batch, True

        def run(self, **params):
            assert params["skip_mode"] == 0 and params["template_end"] == len(matcher._ordered)

        def download(self, stream=0):
            self.busy = False
            return oracle_for(self.batch)

        def close(self):
            pass

    import enzymm_b200.engine as engine_module
    fake_engine = type("E", (), {"compiled": matcher._compile(), "device_library": None, "new_stream": lambda self: 7})()
    monkeypatch.setattr(jess_run.Matcher, "_ensure_engine", lambda self: fake_engine)
    monkeypatch.setattr(engine_module, "Session", FakeSession)
    seen_ids, all_records = [], []
    for ids, headers, records in matcher.scan_files([out], chunk_size=5):
        assert len(ids) == len(headers) <= 5
        seen_ids += ids
        all_records.append(records)
    assert seen_ids == [m.id for m in molecules] and list(offsets.values()) == [0, 5, 10]
    assert 2 <= len(FakeSession.created) <= 3                      # two lanes (one may be re-created to grow)
    assert sum(len(r) for r in all_records) > 10
    text = io.StringIO()
    n_rows = matcher.scan_to_tsv([out], text, chunk_size=len(molecules))
    want = _reference_rows(matcher, _oracle_records(matcher, molecules), molecules)
    assert text.getvalue() == want and n_rows == len(want.splitlines()) - 1 > 5
    # the whole-box call on corpus files: one worker per device shares the chunk plan, chunks come back in input order
    single = list(matcher.scan_files([out], chunk_size=2))
    assert [len(c[0]) for c in single] == [2] * 6
    monkeypatch.setattr(jess_run.Matcher, "close", lambda self: None)
    multi = list(matcher.scan_files([out], chunk_size=2, devices=[0, 1, 2]))
    assert [c[0] for c in multi] == [c[0] for c in single] and all(len(c) == 3 for c in multi)
    assert all(a[2].tobytes() == b[2].tobytes() for a, b in zip(multi, single))
    # two corpus files, a shared queue over the chunk plan (what ranks of a torchrun job would share)
    from enzymm_b200.packing import write_corpus
    second = out.parent / "second.emmpack"
    write_corpus(paths[:5], second, threads=2)
    plan = jess_run.Matcher.corpus_plan([out, second], 4)
    assert plan == [(0, 0, 4, 0), (0, 4, 8, 4), (0, 8, 12, 8), (1, 0, 4, 12), (1, 4, 5, 16)]
    both = list(matcher.scan_files([out, second], chunk_size=4))
    assert [i for c in both for i in c[0]] == [m.id for m in molecules] + [m.id for m in molecules[:5]]
    mine = list(matcher.scan_files([out, second], chunk_size=4, queue=iter([(1, 2), (3, 5)])))
    assert [c[0] for c in mine] == [both[1][0], both[3][0], both[4][0]]
    with pytest.raises(ValueError):
        list(matcher.scan_files([out], devices=[0, 1], queue=iter([(0, 1)])))
