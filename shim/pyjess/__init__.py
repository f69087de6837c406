"""``pyjess`` for an UNMODIFIED EnzyMM: put this directory (and the repository root) in front of
``PYTHONPATH`` and ``import pyjess`` resolves to the B200 engine's stand-in, so the reference's own
modules and command line run on the GPU without a single changed line:

    PYTHONPATH=/path/to/enzymm-b200/shim:/path/to/enzymm-b200 enzymm -i query.pdb -o results.tsv --skip-annotation

(``--skip-annotation`` where the M-CSA annotation blob is not installed.)  The names are the ones EnzyMM
uses (``enzymm/jess_run.py:20``, ``enzymm/template.py:31``); everything lives in
``enzymm_b200.pyjess_api``.  There is no CPU path behind them: without the CUDA library or a GPU the first
query raises.
"""
from enzymm_b200 import __version__  # noqa: F401  (printed by the reference at jess_run.py:659)
from enzymm_b200.pyjess_api import Atom, Hit, Jess, Molecule, Query, Template, TemplateAtom  # noqa: F401

__all__ = ["Atom", "Hit", "Jess", "Molecule", "Query", "Template", "TemplateAtom", "__version__"]
