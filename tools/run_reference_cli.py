"""The reference's unmodified command line (``enzymm._cli.main``) over the ``pyjess`` shim, in this
process: ``python tools/run_reference_cli.py [--device oracle] -- <enzymm arguments>``.

``--device gpu`` (default) is what ``PYTHONPATH=shim:. python -m enzymm ...`` does.  ``--device oracle`` swaps
the shim's device call for the CPU oracle first -- test infrastructure for boxes without a GPU
(``tests/test_reference_own_tests.py``); the product has no such switch."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    argv = sys.argv[1:]
    device = "gpu"
    if argv[:1] == ["--device"]:
        device, argv = argv[1], argv[2:]
    if argv[:1] == ["--"]:
        argv = argv[1:]
    sys.path[:0] = [str(ROOT / "shim"), str(ROOT), str(ROOT / "tests"), str(ROOT / "baseline" / "_ref")]
    blob = ROOT / "baseline" / "_ref" / "enzymm" / "data" / "catalytic_residue_homologs_information.json"
    if blob.parent.is_dir() and not blob.exists():
        blob.write_text("{}")                        # the one blob the reference checkout lacks (.MISSING_LARGE_BLOBS)
    import pyjess
    assert pyjess.Jess.__module__ == "enzymm_b200.pyjess_api", pyjess.__file__
    if device == "oracle":
        from enzymm_b200 import pyjess_api
        from test_reference_dropin import _oracle_device_query
        pyjess_api._device_query = _oracle_device_query
    from enzymm._cli import main as cli
    return cli(argv)


if __name__ == "__main__":
    sys.exit(main())
