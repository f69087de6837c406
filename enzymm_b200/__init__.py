"""enzymm-b200: B200-native geometric template matching behind EnzyMM's Python API.

Drop-in for ONE path of RayHackett/enzymm -- what ``pyjess.Jess(templates).query(...)`` computes
inside ``enzymm.jess_run`` plus EnzyMM's RMSD/orientation filter -- implemented as hand-written
CUDA for sm_100a behind the C ABI in ``include/enzymm_b200.h``.  Module map for a user of the
reference:

    import pyjess                 ->  from enzymm_b200 import pyjess
    from enzymm import template   ->  from enzymm_b200 import template
    from enzymm import jess_run   ->  from enzymm_b200 import jess_run
"""
__version__ = "0.1.0"

from .structures import Atom, Molecule  # noqa: F401,E402
from .template_atoms import TemplateAtom, JessTemplate  # noqa: F401,E402
from . import templates as template  # noqa: F401,E402
from . import pyjess_api as pyjess  # noqa: F401,E402
from . import matcher as jess_run  # noqa: F401,E402
from .matcher import Match, Matcher, load_molecules  # noqa: F401,E402
from .templates import Template, load_templates  # noqa: F401,E402
