"""EnzyMM's template model: ``Vec3``, ``Residue``, ``Cluster``, ``Template``, ``load_templates``.

Only what the matching hot path needs is carried over (§8a rows a4,
a5): the residue triplets, the per-residue orientation definition that the fused filter kernel
consumes (``enzymm/template.py:213-304``), ``effective_size`` (which selects the per-size
thresholds and the logistic models, ``template.py:786-808``) and the REMARK metadata that the
result rows print.  M-CSA residue-level annotation (``AnnotatedTemplate.derive_mcsa_annotations``)
is OUT OF SCOPE: its data blob is not shipped with the reference checkout
(``.MISSING_LARGE_BLOBS``), so ``AnnotatedTemplate.load`` degrades to a plain ``Template``.
"""
from __future__ import annotations

import gzip
import io
import json
import math
import os
import re
import warnings
from dataclasses import dataclass
from functools import cached_property
from pathlib import Path
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, TextIO, Tuple, Union

from .template_atoms import JessTemplate, TemplateAtom

__all__ = ["Vec3", "Residue", "Cluster", "Template", "AnnotatedTemplate", "load_templates",
           "check_template", "ORIENTATION_ATOMS", "rank_order"]

_DATA = Path(__file__).resolve().parent / "data"
_BUNDLE = _DATA / "jess_templates_20230210.bundle.gz"


def rank_order(values: Sequence[int]) -> List[int]:
    """Dense 1-based ranks, e.g. ``[40, 10, 40, 7] -> [3, 2, 3, 1]`` (``enzymm/utils.py:10-22``)."""
    position = {v: i + 1 for i, v in enumerate(sorted(set(values)))}
    return [position[v] for v in values]


@dataclass(frozen=True)
class Vec3:
    """3-vector in double precision (``enzymm/template.py:43-181``)."""

    x: float
    y: float
    z: float

    def __post_init__(self):
        if self.x != self.x or self.y != self.y or self.z != self.z:
            raise ValueError("Cannot create a Vec3 with NaN values. Likely the Jess superposition failed.")

    @classmethod
    def from_xyz(cls, item) -> "Vec3":
        return cls(item.x, item.y, item.z)

    @property
    def norm(self) -> float:
        return math.sqrt(self.x**2 + self.y**2 + self.z**2)

    def normalize(self) -> "Vec3":
        n = self.norm
        return Vec3(self.x, self.y, self.z) if n == 0 else self / n

    def _lift(self, other, what: str) -> Tuple[float, float, float]:
        if isinstance(other, Vec3):
            return other.x, other.y, other.z
        if isinstance(other, (int, float)) and not isinstance(other, bool):
            return other, other, other
        raise TypeError(f"Expected int, float or Vec3, got {type(other).__name__}")

    def __matmul__(self, other: "Vec3") -> float:
        if not isinstance(other, Vec3):
            raise TypeError(f"Expected Vec3, got {type(other).__name__}")
        return self.x * other.x + self.y * other.y + self.z * other.z

    def __add__(self, other) -> "Vec3":
        a, b, c = self._lift(other, "+")
        return Vec3(self.x + a, self.y + b, self.z + c)

    def __sub__(self, other) -> "Vec3":
        a, b, c = self._lift(other, "-")
        return Vec3(self.x - a, self.y - b, self.z - c)

    def __truediv__(self, other) -> "Vec3":
        a, b, c = self._lift(other, "/")
        return Vec3(self.x / a, self.y / b, self.z / c)

    def angle_to(self, other: "Vec3") -> float:
        """Angle in radians; cosines a hair outside [-1, 1] snap to 0 / pi
        (``enzymm/template.py:157-181``)."""
        cosine = self.normalize() @ other.normalize()
        if -1 <= cosine <= 1:
            return math.acos(cosine)
        if math.isclose(cosine, 1, rel_tol=1e-5):
            return 0
        if math.isclose(cosine, -1, rel_tol=1e-5):
            return math.pi
        raise ValueError(
            f"ArcCos is not defined outside [-1,1]. self.vec is {[self.x, self.y, self.z]}, "
            f"other vec is {[other.x, other.y, other.z]}")


# residue name -> (tail atom, head atom | "mid"): the axis a residue's orientation is measured on.
# "mid" = midpoint of the two atoms that are not the tail (symmetric groups).
# Restates the table at enzymm/template.py:233-256.
ORIENTATION_ATOMS: Dict[str, Tuple[str, str]] = {
    "GLY": ("C", "O"), "PRO": ("C", "O"), "ANY": ("C", "O"),
    "ALA": ("CA", "CB"), "VAL": ("CA", "CB"), "LEU": ("CA", "CB"), "ILE": ("CA", "CB"),
    "PTM": ("CA", "CB"),
    "MET": ("CG", "SD"), "CYS": ("CB", "SG"), "SER": ("CB", "OG"), "THR": ("CB", "OG1"),
    "TYR": ("CZ", "OH"), "TRP": ("CZ2", "NE1"), "ASN": ("CG", "OD1"), "GLN": ("CD", "OE1"),
    "LYS": ("CE", "NZ"), "HIS": ("CG", "ND1"),
    "PHE": ("CZ", "mid"), "ARG": ("CZ", "mid"), "ASP": ("CG", "mid"), "GLU": ("CD", "mid"),
}

MIDPOINT = 9  # sentinel second index meaning "midpoint of the other two atoms"


class Residue:
    """Three template atoms of one residue + its orientation axis (``enzymm/template.py:184-422``)."""

    __slots__ = ("_atoms", "_vec", "_indices")

    def __init__(self, atoms: Sequence[TemplateAtom] = None, *, _atoms=None, _vec=None, _indices=None):
        if atoms is None:  # field-wise construction (used by annotated subclasses upstream)
            atoms, vec, indices = tuple(_atoms), _vec, _indices
        else:
            atoms = tuple(atoms)
            vec, indices = self.calc_residue_orientation(atoms)
        object.__setattr__(self, "_atoms", atoms)
        object.__setattr__(self, "_vec", vec)
        object.__setattr__(self, "_indices", indices)

    def __setattr__(self, key, value):
        raise AttributeError("Residue is immutable")

    @staticmethod
    def calc_residue_orientation(atoms: Sequence[TemplateAtom]) -> Tuple[Vec3, Tuple[int, int]]:
        """Orientation vector and the atom index pair it is taken between.

        ``(i, j)`` -> vector from atom ``i`` to atom ``j``; ``(i, 9)`` -> from atom ``i`` to the
        midpoint of the other two (``enzymm/template.py:213-304``).  ``KeyError`` for a residue
        name without a definition, ``ValueError`` if a named atom is absent.
        """
        resname = atoms[0].residue_names[0]
        try:
            tail_name, head_name = ORIENTATION_ATOMS[resname]
        except KeyError as exc:
            raise KeyError(f"Residue orientation is not defined for the residue type {resname}") from exc

        def find(name: str, role: str) -> int:
            for i, atom in enumerate(atoms):
                if atom.atom_names[0] == name:
                    return i
            raise ValueError(f"Failed to find {role} atom for amino-acid {resname!r}")

        if head_name == "mid":
            tail = find(tail_name, "middle")
            others = [a for a in atoms if a != atoms[tail]]
            if len(others) != 2:
                raise ValueError(f"Failed to find two side atoms for amino-acid {resname!r}")
            mid = (Vec3.from_xyz(others[0]) + Vec3.from_xyz(others[1])) / 2
            return mid - Vec3.from_xyz(atoms[tail]), (tail, MIDPOINT)
        tail = find(tail_name, "first")
        head = find(head_name, "second")
        return Vec3.from_xyz(atoms[head]) - Vec3.from_xyz(atoms[tail]), (tail, head)

    @classmethod
    def construct_residues_from_atoms(cls, atoms: Iterable[TemplateAtom]) -> List["Residue"]:
        atoms = list(atoms)
        residues = []
        for start in range(0, len(atoms), 3):
            triplet = tuple(atoms[start:start + 3])
            if len(triplet) != 3:
                raise ValueError(f"Failed to construct residues. Got only {len(triplet)} ATOM lines")
            owners = {(a.residue_names[0], a.chain_id, a.residue_number) for a in triplet}
            if len(owners) != 1:
                raise ValueError(
                    "Failed to construct residues. Atoms of different match_mode, chains, residue "
                    f"types or residue numbers {owners} found in Atom triplet")
            residues.append(cls(triplet))
        return residues

    atoms = property(lambda self: self._atoms)
    residue_name = property(lambda self: self._atoms[0].residue_names[0])
    match_mode = property(lambda self: self._atoms[0].match_mode)
    residue_number = property(lambda self: self._atoms[0].residue_number)
    chain_id = property(lambda self: self._atoms[0].chain_id)
    orientation_vector = property(lambda self: self._vec)
    orientation_vector_indices = property(lambda self: self._indices)

    @property
    def allowed_residues(self) -> str:
        from .template_atoms import ONE_TO_THREE
        three_to_one = {v: k for k, v in ONE_TO_THREE.items() if k != "U"}
        return "".join({three_to_one[r] for r in self._atoms[0].residue_names})

    @property
    def backbone(self) -> bool:
        names = self._atoms[0].residue_names
        return "ANY" in names or "XXX" in names

    def __eq__(self, other):
        if not isinstance(other, Residue):
            return NotImplemented
        return (self._atoms, self._vec, self._indices) == (other._atoms, other._vec, other._indices)

    def __hash__(self):
        return hash((self._atoms, self._vec, self._indices))

    def __repr__(self):
        return f"Residue({self.residue_name} {self.chain_id}{self.residue_number})"


@dataclass(frozen=True)
class Cluster:
    """``REMARK CLUSTER <id>_<member>_<size>`` (``enzymm/template.py:435-452``)."""

    id: int
    member: int
    size: int

    def __post_init__(self):
        if self.member > self.size:
            raise ValueError("Cluster member cannot be greater than cluster size")


def _load_json(name: str):
    with open(_DATA / name) as handle:
        return json.load(handle)


class Template(JessTemplate):
    """A catalytic-site template: residue triplets + REMARK metadata (``enzymm/template.py:455``)."""

    _CATH_MAPPING: Dict[str, List[str]] = _load_json("MCSA_CATH_mapping.json")
    _EC_MAPPING: Dict[str, str] = _load_json("MCSA_EC_mapping.json")
    _PDB_SIFTS: Dict[str, dict] = _load_json("pdb_sifts.json")

    _META_FIELDS = ("pdb_id", "mcsa_id", "template_id_string", "cluster", "uniprot_id", "organism",
                    "organism_id", "resolution", "experimental_method", "enzyme_discription",
                    "represented_sites")

    def __init__(self, residues: Sequence[Residue], pdb_id: Optional[str] = None,
                 mcsa_id: Optional[int] = None, *, id: Optional[str] = None,
                 template_id_string: Optional[str] = None, cluster: Optional[Cluster] = None,
                 uniprot_id: Optional[str] = None, organism: Optional[str] = None,
                 organism_id: Optional[str] = None, resolution: Optional[float] = None,
                 experimental_method: Optional[str] = None, enzyme_discription: Optional[str] = None,
                 represented_sites: Optional[int] = None, ec: Iterable[str] = (),
                 cath: Iterable[str] = ()):
        if len(residues) == 0:
            raise ValueError("Tried creating an `Template` from an empty list of residues!")
        kind = type(residues[0])
        if any(not isinstance(r, kind) for r in residues):
            raise ValueError("Tried creating a `Template` from different types of `Residue` objects.")
        super().__init__([a for r in residues for a in r.atoms], id=id)
        self.residues = tuple(residues)
        self.pdb_id = pdb_id
        self.mcsa_id = mcsa_id
        self.template_id_string = template_id_string
        self.cluster = cluster
        self.uniprot_id = uniprot_id
        self.organism = organism
        self.organism_id = organism_id
        self.resolution = resolution
        self.experimental_method = experimental_method
        self.enzyme_discription = enzyme_discription
        self.represented_sites = represented_sites
        self.ec = tuple(sorted({*ec, *self._mapped_ec()}))
        self.cath = tuple(sorted({*cath, *self._mapped_cath()}))

    # -- derived properties -------------------------------------------------------------------
    @cached_property
    def effective_size(self) -> int:
        """Residues that are residue-type specific (mode < 100) and not backbone wildcards
        (``enzymm/template.py:786-808``).  Selects thresholds and logistic models."""
        return sum(1 for r in self.residues if r.match_mode < 100 and not r.backbone)

    @cached_property
    def multimeric(self) -> bool:
        first = self.residues[0].chain_id
        return any(r.chain_id != first for r in self.residues)

    @cached_property
    def relative_order(self) -> List[int]:
        if self.multimeric:
            return [0]
        return rank_order([r.residue_number for r in self.residues])

    def _pdb_chains(self):
        return {f"{self.pdb_id}{r.chain_id}" for r in self.residues}

    def _sifts(self, field: str) -> List[str]:
        found: List[str] = []
        if self.pdb_id is not None:
            for chain in self._pdb_chains():
                entry = self._PDB_SIFTS.get(chain)
                if entry is not None:
                    found.extend(v for v in entry.get(field, ()) if v != "?")
        return found

    def _mapped_cath(self) -> List[str]:
        found: List[str] = []
        if self.mcsa_id:
            found.extend(self._CATH_MAPPING[str(self.mcsa_id)])
        if self.pdb_id:
            found.extend(self._sifts("cath"))
        return found

    def _mapped_ec(self) -> List[str]:
        found: List[str] = []
        if self.mcsa_id is not None:
            found.append(self._EC_MAPPING[str(self.mcsa_id)])
        found.extend(self._sifts("ec"))
        return found

    # -- identity (Matcher rejects duplicate templates via set(), jess_run.py:643-646) -----------
    def _state(self) -> Tuple:
        plain = tuple(Residue(r.atoms) for r in self.residues)
        return (plain, self.id, self.pdb_id, self.mcsa_id, self.template_id_string, self.cluster,
                self.effective_size, self.uniprot_id, self.organism, self.organism_id,
                self.resolution, self.experimental_method, self.enzyme_discription,
                self.represented_sites, self.ec, self.cath)

    def __eq__(self, other):
        if not isinstance(other, Template):
            return NotImplemented
        return self._state() == other._state()

    def __ne__(self, other):
        if not isinstance(other, Template):
            return NotImplemented
        return self._state() != other._state()

    def __hash__(self):
        return hash((type(self), self._state()))

    def copy(self) -> "Template":
        meta = {k: getattr(self, k) for k in self._META_FIELDS}
        return type(self)(residues=self.residues, id=self.id, ec=self.ec, cath=self.cath, **meta)

    __copy__ = copy

    # -- parsing ------------------------------------------------------------------------------
    @classmethod
    def loads(cls, text: str, id: Optional[str] = None, warn: bool = False, **kw) -> "Template":
        return cls.load(io.StringIO(text), id=id, warn=warn, **kw)

    @classmethod
    def load(cls, file: Union[TextIO, Iterator[str], str, os.PathLike], id: Optional[str] = None,
             warn: bool = False, **_ignored) -> "Template":
        """Parse one template file: REMARK metadata + ATOM triplets (``enzymm/template.py:611-693``).
        ``HETATM`` records and duplicated ATOM lines are ``ValueError``s."""
        if isinstance(file, (str, os.PathLike)):
            with open(os.fspath(file)) as handle:
                lines = handle.readlines()
        else:
            lines = list(file)
        meta: Dict[str, object] = {"ec": [], "cath": []}
        atoms: List[TemplateAtom] = []
        seen = set()
        for line in lines:
            tokens = line.split()
            if not tokens:
                continue
            tag = tokens[0]
            if tag == "REMARK":
                if len(tokens) == 1:
                    continue
                handler = _REMARK_HANDLERS.get(tokens[1])
                if handler is None:
                    continue
                if len(tokens) < 3:
                    raise IndexError(f"Expected some annotation after the REMARK {tokens[1]} flag")
                if tokens[2].upper() in ("NONE", "?", "NAN", "NA"):
                    continue
                handler(tokens, meta, warn)
            elif tag == "ATOM":
                if line in seen:
                    raise ValueError("Duplicate Atom lines passed!")
                seen.add(line)
                atoms.append(TemplateAtom.loads(line))
            elif tag == "HETATM":
                raise ValueError("Supplied template with HETATM record. HETATMs cannot be searched by Jess")
        residues = Residue.construct_residues_from_atoms(atoms)
        return Template(residues=residues, id=id, **meta)  # type: ignore[arg-type]


class AnnotatedTemplate(Template):
    """Upstream adds per-residue M-CSA annotation here (``enzymm/template.py:1024-1456``); the
    annotation blob is absent from the reference checkout, so loading yields a plain ``Template``."""

    @classmethod
    def load(cls, file, id: Optional[str] = None, warn: bool = False,
             with_annotations: bool = True) -> Template:
        return Template.load(file, id=id, warn=warn)

    @classmethod
    def loads(cls, text: str, id: Optional[str] = None, warn: bool = False,
              with_annotations: bool = True) -> Template:
        return Template.load(io.StringIO(text), id=id, warn=warn)


# ---- REMARK handlers (enzymm/template.py:887-1021) ---------------------------------------------
_UNIPROT = re.compile(r"[OPQ][0-9][A-Z0-9]{3}[0-9]|[A-NR-Z][0-9]([A-Z][A-Z0-9]{2}[0-9]){1,2}")
_EC_STRICT = re.compile(r"[1-7](\.(\-|\d{1,})){3}")
_EC_LOOSE = re.compile(r"[1-7](\.(\-|\d{1,}|n\d{1,})){3}")
_CATH = re.compile(r"[1-46](\.(\-|\d{1,})){3}")


def _r_pdb(tokens, meta, warn):
    if len(tokens[2]) != 4:
        raise ValueError(f"Found {tokens[2]} which has more than the expected 4 characters of a PDB_ID.")
    meta["pdb_id"] = tokens[2].lower()


def _r_uniprot(tokens, meta, warn):
    found = _UNIPROT.search(tokens[2])
    if not found:
        raise ValueError(f"Did not find a valid UniProt ID, found {tokens[2]}")
    meta["uniprot_id"] = found.group()


def _r_int(field, what):
    def handler(tokens, meta, warn):
        try:
            meta[field] = int(tokens[2])
        except ValueError as exc:
            raise ValueError(f"Did not find {what}, found {tokens[2]}") from exc
    return handler


def _r_cluster(tokens, meta, warn):
    try:
        meta["cluster"] = Cluster(*(int(v) for v in tokens[2].split("_")))
    except (ValueError, TypeError) as exc:
        raise ValueError(
            f"Did not find a Cluster specification in the form <id>_<member>_<size>, found {tokens[2]}") from exc


def _r_resolution(tokens, meta, warn):
    try:
        meta["resolution"] = float(tokens[2])
    except ValueError as exc:
        raise ValueError(f"Ill-formatted pdb resolution: {tokens[2]}") from exc


def _r_words(field):
    def handler(tokens, meta, warn):
        meta[field] = " ".join(tokens[2:])
    return handler


def _r_token(field):
    def handler(tokens, meta, warn):
        meta[field] = tokens[2]
    return handler


def _extend_unique(target: list, values):
    for v in values:
        if v not in target:
            target.append(v)


def _r_ec(tokens, meta, warn):
    strict = [m.group() for m in _EC_STRICT.finditer(tokens[2])]
    loose = [m.group() for m in _EC_LOOSE.finditer(tokens[2])]
    if strict:
        _extend_unique(meta["ec"], strict)
    elif loose:
        _extend_unique(meta["ec"], loose)
        if warn:
            warnings.warn(f"Rare EC number(s) {loose} presumed to be noncatalytic detected!")
    else:
        raise ValueError(f"Did not find a valid EC number, found {tokens[2]}")


def _r_cath(tokens, meta, warn):
    found = [m.group() for m in _CATH.finditer(tokens[2])]
    if not found:
        raise ValueError(f"Did not find a valid CATH number, found {tokens[2]}")
    _extend_unique(meta["cath"], found)


_REMARK_HANDLERS = {
    "ID": _r_token("template_id_string"),
    "PDB_ID": _r_pdb,
    "UNIPROT_ID": _r_uniprot,
    "MCSA_ID": _r_int("mcsa_id", "a M-CSA ID"),
    "CLUSTER": _r_cluster,
    "ORGANISM_NAME": _r_words("organism"),
    "ORGANISM_ID": _r_token("organism_id"),
    "RESOLUTION": _r_resolution,
    "EXPERIMENTAL_METHOD": _r_words("experimental_method"),
    "EC": _r_ec,
    "CATH": _r_cath,
    "ENZYME": _r_words("enzyme_discription"),
    "REPRESENTING": _r_int("represented_sites", "a number of represented sites"),
}


# ---- library loading ----------------------------------------------------------------------------
def iter_bundle(bundle: Path = _BUNDLE) -> Iterator[Tuple[str, str]]:
    """Yield ``(relative_path, text)`` for every template in the packed shipped library."""
    with gzip.open(bundle, "rt") as handle:
        name, chunk = None, []
        for line in handle:
            if line.startswith("@@ "):
                if name is not None:
                    yield name, "".join(chunk)
                name, chunk = line[3:].strip(), []
            else:
                chunk.append(line)
        if name is not None:
            yield name, "".join(chunk)


def load_templates(template_dir: Optional[Path] = None, warn: bool = False, verbose: bool = False,
                   cpus: int = 0, with_annotations: bool = True,
                   subset: Optional[str] = None) -> Iterator[Template]:
    """Yield every template below ``template_dir`` (recursive ``*.pdb``), or the shipped
    jess_templates_20230210 library when ``None`` (``enzymm/template.py:1505-1575``).

    ``subset`` restricts the shipped library to relative paths starting with that prefix.
    ``NotADirectoryError`` / ``FileNotFoundError`` / ``ValueError`` as upstream.
    """
    if template_dir is None:
        for name, text in iter_bundle():
            if subset is not None and not name.startswith(subset):
                continue
            try:
                yield AnnotatedTemplate.loads(text, warn=warn, with_annotations=with_annotations)
            except (ValueError, KeyError) as exc:
                raise ValueError(f"Shipped template {name} could not be parsed") from exc
        return
    template_dir = Path(template_dir)
    if not template_dir.is_dir():
        raise NotADirectoryError(f"The path {template_dir} doesnt exist or is not a directory!")
    if verbose:
        print(f"Loading Template files from {template_dir.resolve()}")
    paths = sorted(template_dir.glob("**/*.pdb"))
    if not paths:
        raise FileNotFoundError(
            f"No template files with the .pdb extension found in the {template_dir.resolve()} directory")
    for path in paths:
        try:
            yield AnnotatedTemplate.load(path, warn=warn, with_annotations=with_annotations)
        except ValueError as exc:
            raise ValueError(
                f"Passed Template file {path.resolve()} contained ATOM lines which are not in Jess Template format.") from exc
        except KeyError as exc:
            raise ValueError(f"Passed Template file {path.resolve()} contained issues with some residues.") from exc


def check_template(template: Template, warn: bool = True) -> bool:
    """Annotation sanity check; always ``True`` with ``warn=False`` (``enzymm/template.py:1579-1618``)."""
    if not warn:
        return True
    ok = True
    if not template.ec:
        ok = False
        warnings.warn("Could not find EC number annotations")
    if not template.cath:
        ok = False
        warnings.warn("Could not find CATH annotations")
    if template.pdb_id:
        for chain in template._pdb_chains():
            entry = template._PDB_SIFTS.get(chain)
            if entry and template.uniprot_id != entry["uniprot_id"]:
                ok = False
                warnings.warn(f"Different UniProt Accessions {template.uniprot_id} and {entry['uniprot_id']} found")
    return ok
