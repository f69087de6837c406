"""The reference's UNMODIFIED command line on the GPU engine, switched over by ``PYTHONPATH`` alone:
``PYTHONPATH=shim:.:baseline/_ref python -m enzymm ...`` resolves ``import pyjess`` to ``shim/pyjess`` (this
repo's stand-in) and must write, after its ``# Version`` line, the table ``Matcher.scan_to_tsv`` writes for
the same files and templates.  ``tests/test_reference_own_tests.py`` checks the same equality with the
oracle in place of the device; this file sorts last in the ``-m gpu`` run and had its first run on the
driver's box (written after the round's GPU budget was spent)."""
import glob
import io
import os
import shutil
import subprocess
import sys
from pathlib import Path

import pytest

from conftest import GOLDEN, ROOT
from enzymm_b200 import jess_run, template

pytestmark = pytest.mark.gpu
REF = ROOT / "baseline" / "_ref" / "enzymm"


def test_reference_cli_by_pythonpath_equals_scan_to_tsv(tmp_path):
    if not (REF / "jess_run.py").exists():
        pytest.skip("baseline/_ref is absent: run __graft_entry__.build() where /root/reference exists")
    blob = REF / "data" / "catalytic_residue_homologs_information.json"
    if not blob.exists():
        blob.write_text("{}")                        # the one blob the reference checkout lacks (.MISSING_LARGE_BLOBS)
    tdir = tmp_path / "templates"
    for size in ("3_residues", "4_residues", "5_residues"):
        for entry in ("csa3d_0285", "csa3d_0045", "csa3d_0421", "csa3d_0415"):
            src = REF / "jess_templates_20230210" / size / "results" / entry
            if src.is_dir():
                shutil.copytree(src, tdir / size / "results" / entry)
    paths = [GOLDEN / "1AMY.pdb", GOLDEN / "AF-P0DUB6-F1-model_v4.pdb"]
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([str(ROOT / "shim"), str(ROOT), str(ROOT / "baseline" / "_ref")]))
    # the reference lists the directory with an unsorted recursive glob (template.py:1490-1495); same list here
    templates = [template.AnnotatedTemplate.load(Path(p), warn=False, with_annotations=False)
                 for p in glob.glob(f"{tdir}/**/*.pdb", recursive=True)]
    for extra in ([], ["--unfiltered"], ["--skip-smaller-hits"]):
        out = tmp_path / "cli.tsv"
        run = subprocess.run([sys.executable, "-m", "enzymm", "-i", str(paths[0]), "-i", str(paths[1]), "-o", str(out),
                              "-t", str(tdir), "--skip-annotation", "-n", "8"] + extra,
                             capture_output=True, text=True, cwd=tmp_path, env=env, timeout=900)
        assert run.returncode == 0, (run.stdout + run.stderr)[-3000:]
        version, _, table = out.read_text().partition("\n")
        assert version.startswith("# Version")
        matcher = jess_run.Matcher(templates, filter_matches="--unfiltered" not in extra,
                                   skip_smaller_hits="--skip-smaller-hits" in extra)
        try:
            mine = io.StringIO()
            rows = matcher.scan_to_tsv(paths, mine, predict_correctness="--unfiltered" not in extra)
        finally:
            matcher.close()
        assert rows == len(table.splitlines()) - 1 > 4
        assert mine.getvalue() == table, extra
