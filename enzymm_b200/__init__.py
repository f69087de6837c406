"""enzymm-b200: B200-native geometric template matching behind EnzyMM's Python API."""
__version__ = "0.1.0"

from .structures import Atom, Molecule  # noqa: F401
from .template_atoms import TemplateAtom, JessTemplate  # noqa: F401
