"""Run the reference's OWN, unmodified test files (``/root/reference/tests``) on top of this repo's
``pyjess`` shim and print one line per test.

The reference package is the unmodified install under ``baseline/_ref`` (``__graft_entry__.build()``),
``sys.modules["pyjess"]`` is ``enzymm_b200.pyjess``.  ``--device oracle`` replaces the device call of the
shim with the CPU oracle (test infrastructure; what a box without a GPU can do), ``--device gpu`` leaves
the CUDA path in place.  Needs ``/root/reference`` for the test files themselves -- they are reference
content and are not copied into this repository -- so it runs in the build container only.  Run it from
``/root/reference``: the reference's CLI test reads ``tests/test_data/input_list.txt``, whose entries are
relative to that directory.

``--annotations placeholder``: the reference's ``AnnotatedTemplate`` needs
``data/catalytic_residue_homologs_information.json`` (M-CSA homolog data), which its checkout lacks
(``.MISSING_LARGE_BLOBS``); with the default ``{}`` stub every test that touches an annotated template stops
inside the reference before reaching ``pyjess``.  The placeholder is a file of the same SHAPE generated
from the template library itself -- every template residue listed as its own reference residue, no roles,
no PTMs, assembly 1 -- written next to a temporary COPY of the installed package.  It is not M-CSA data:
the annotation columns it yields mean nothing (and the reference tests that assert them fail, as they
should); what it buys is that ``TestMatch`` / ``TestMatcher`` / the CLI test run their hot-path assertions
unmodified over the shim.

usage: python tools/run_reference_tests.py [--device oracle|gpu] [--annotations stub|placeholder] [--json out.json] [pattern ...]
"""
import argparse
import importlib
import json
import sys
import unittest
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF_PACKAGE = ROOT / "baseline" / "_ref"
REF_CHECKOUT = Path("/root/reference")


def placeholder_blob() -> dict:
    """{mcsa_id: {pdb_id + chain: HomologousPDB fields}} covering every residue of every shipped template
    (``enzymm/mcsa_info.py:90-148`` for the shape, ``template.py:1310-1345`` for the look-ups)."""
    from enzymm_b200.templates import load_templates
    data: dict = {}
    for t in load_templates():
        if t.mcsa_id is None or t.pdb_id is None:
            continue
        entry = data.setdefault(str(int(t.mcsa_id)), {})
        for residue in t.residues:
            for chain in {residue.chain_id, residue.chain_id[:1]}:
                key = t.pdb_id.lower() + chain
                pdb = entry.setdefault(key, {"reference_pdbchain": key, "is_reference": True, "pdb_id": t.pdb_id.lower(),
                                             "chain_name": chain, "assembly_chain_name": chain, "assembly": 1,
                                             "residues": {}})
                known = {r["auth_resid"] for r in pdb["residues"].values()}
                if residue.residue_number not in known:
                    pdb["residues"][str(len(pdb["residues"]))] = {
                        "code": residue.residue_name, "resid": residue.residue_number, "auth_resid": residue.residue_number,
                        "function_location_abv": None, "ptm": None, "roles": [], "roles_summary": []}
    return data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", choices=("oracle", "gpu"), default="oracle")
    ap.add_argument("--json", default=None)
    ap.add_argument("--annotations", choices=("stub", "placeholder"), default="stub")
    ap.add_argument("patterns", nargs="*", default=["test_jess_run", "test_template", "test_utils", "test_cli"])
    args = ap.parse_args()
    if not (REF_PACKAGE / "enzymm" / "jess_run.py").exists() or not (REF_CHECKOUT / "tests").is_dir():
        print("reference package or checkout absent", file=sys.stderr)
        return 3
    blob = REF_PACKAGE / "enzymm" / "data" / "catalytic_residue_homologs_information.json"
    if not blob.exists():
        blob.write_text("{}")                        # the one blob the reference checkout lacks (.MISSING_LARGE_BLOBS)
    package_root = REF_PACKAGE
    sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
    if args.annotations == "placeholder":
        import shutil
        import tempfile
        package_root = Path(tempfile.mkdtemp(prefix="emm_ref_"))
        shutil.copytree(REF_PACKAGE / "enzymm", package_root / "enzymm")
        (package_root / "enzymm" / "data" / "catalytic_residue_homologs_information.json").write_text(json.dumps(placeholder_blob()))
    sys.path[2:2] = [str(package_root), str(REF_CHECKOUT)]
    from enzymm_b200 import pyjess as shim
    sys.modules["pyjess"] = shim
    if args.device == "oracle":
        from enzymm_b200 import pyjess_api
        from test_reference_dropin import _oracle_device_query
        pyjess_api._device_query = _oracle_device_query
    enzymm = importlib.import_module("enzymm")
    assert str(package_root) in enzymm.__file__, enzymm.__file__
    suite = unittest.TestSuite()
    loader = unittest.TestLoader()
    for name in args.patterns:
        suite.addTests(loader.loadTestsFromName(f"tests.{name}"))

    class Result(unittest.TextTestResult):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self.outcomes = {}

        def addSuccess(self, test):
            super().addSuccess(test)
            self.outcomes[test.id()] = ("pass", "")

        def addFailure(self, test, err):
            super().addFailure(test, err)
            self.outcomes[test.id()] = ("fail", str(err[1])[:4000])

        def addError(self, test, err):
            super().addError(test, err)
            self.outcomes[test.id()] = ("error", f"{err[0].__name__}: {err[1]}"[:400])

        def addSkip(self, test, reason):
            super().addSkip(test, reason)
            self.outcomes[test.id()] = ("skip", reason)

        def addExpectedFailure(self, test, err):
            super().addExpectedFailure(test, err)
            self.outcomes[test.id()] = ("xfail", "")

    runner = unittest.TextTestRunner(resultclass=Result, verbosity=0, stream=open("/dev/null", "w"))
    result = runner.run(suite)
    for test_id, (outcome, why) in sorted(result.outcomes.items()):
        print(f"{outcome:6s} {test_id}" + (f"    {why.splitlines()[0] if why else ''}" if outcome not in ("pass",) else ""))
    counts = {}
    for outcome, _ in result.outcomes.values():
        counts[outcome] = counts.get(outcome, 0) + 1
    print("TOTAL", json.dumps(counts, sort_keys=True))
    if args.json:
        Path(args.json).write_text(json.dumps({k: list(v) for k, v in result.outcomes.items()}, indent=1, sort_keys=True))
    if package_root != REF_PACKAGE:
        import shutil
        shutil.rmtree(package_root, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
