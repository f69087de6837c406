"""GPU leg of the packed corpus files (``tests/test_corpus.py`` holds the host logic): screening a
``*.emmpack`` file gives the records and the table that screening the text files it was written from
gives.  Runs after the parity tests proper (file name sorts last): the feature was written after the
round's GPU budget was spent, so this test's first run is the driver's."""
import io

import pytest

from conftest import GOLDEN
from enzymm_b200 import jess_run
from enzymm_b200.packing import write_corpus
from enzymm_b200.synth import SynthConfig, generate_chunk

pytestmark = pytest.mark.gpu


def test_scan_corpus_equals_scan_files(tmp_path, active_templates):
    chunk = generate_chunk(21, SynthConfig(n_residues=180), templates=active_templates, count=6)
    paths = []
    for i in range(chunk.n_structures):
        p = tmp_path / f"synth{i}.pdb"
        p.write_text(chunk.to_pdb(i))
        paths.append(p)
    paths[3:3] = [GOLDEN / "1AMY.pdb", GOLDEN / "AF-P0DUB6-F1-model_v4.pdb"]
    corpus = tmp_path / "all.emmpack"
    assert write_corpus(paths, corpus, threads=2) == len(paths)
    for kwargs in (dict(), dict(filter_matches=False, skip_smaller_hits=True)):
        matcher = jess_run.Matcher(templates=active_templates, **kwargs)
        try:
            from_text = list(matcher.scan_files(paths, chunk_size=3, threads=2))
            from_corpus = list(matcher.scan_files([corpus], chunk_size=3))
            assert len(from_text) == len(from_corpus) == 3
            for (chunk_paths, _, want), (ids, _, got) in zip(from_text, from_corpus):
                assert ids == [p.stem for p in map(type(paths[0]), chunk_paths)]
                assert got.tobytes() == want.tobytes()
            assert sum(len(r) for _, _, r in from_corpus) > 10
            a, b = io.StringIO(), io.StringIO()
            assert matcher.scan_to_tsv(paths, a, chunk_size=4, threads=2) == matcher.scan_to_tsv([corpus], b, chunk_size=4) > 0
            assert a.getvalue() == b.getvalue()
            # the whole-box call on the corpus: two workers (sharing device 0 on a one-GPU box) pull from one chunk plan
            shared = list(matcher.scan_files([corpus], chunk_size=3, devices=[0, 0]))
            assert [c[0] for c in shared] == [c[0] for c in from_corpus]
            assert all(x[2].tobytes() == y[2].tobytes() for x, y in zip(shared, from_corpus))
        finally:
            matcher.close()
