"""Template-side pyjess surface: ``TemplateAtom`` and the subclassable ``Template`` base.

Host-side library-compile stage of the hot path.  The reference gets
both types from ``pyjess`` (call sites ``enzymm/template.py:679`` ``TemplateAtom.loads`` and
``template.py:455,535`` ``class Template(pyjess.Template)`` / ``super().__init__(atoms, id=id)``).

Template ATOM line layout, established from all 98 808 lines of the shipped library
(SURVEY.md §8a "Template ATOM line format"): cols 1-6 ``ATOM``, 7-11 match_mode, 13-16 atom
name, 18-20 residue name, 21-22 chain id, 23-26 residue number, 31-54 x/y/z, 55-60 alternate
residue one-letter codes, 61-66 distance weight.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Sequence, Tuple

__all__ = ["TemplateAtom", "JessTemplate", "ONE_TO_THREE"]

ONE_TO_THREE = {
    "A": "ALA", "C": "CYS", "D": "ASP", "E": "GLU", "F": "PHE", "G": "GLY", "H": "HIS",
    "I": "ILE", "K": "LYS", "L": "LEU", "M": "MET", "N": "ASN", "P": "PRO", "Q": "GLN",
    "R": "ARG", "S": "SER", "T": "THR", "V": "VAL", "W": "TRP", "Y": "TYR",
    "X": "XXX", "U": "SEC",
}


class TemplateAtom:
    """One template atom = one typing predicate + one point (mirror of ``pyjess.TemplateAtom``;
    ctor keywords as at ``tests/test_template.py:234-268``)."""

    __slots__ = ("chain_id", "residue_number", "residue_names", "atom_names",
                 "distance_weight", "match_mode", "x", "y", "z")

    def __init__(self, *, chain_id: str, residue_number: int, residue_names: Sequence[str],
                 atom_names: Sequence[str], x: float, y: float, z: float,
                 distance_weight: float = 0.0, match_mode: int = 0):
        if not residue_names:
            raise ValueError("TemplateAtom needs at least one residue name")
        if not atom_names:
            raise ValueError("TemplateAtom needs at least one atom name")
        self.chain_id = str(chain_id)
        self.residue_number = int(residue_number)
        self.residue_names = tuple(str(r) for r in residue_names)
        self.atom_names = tuple(str(a) for a in atom_names)
        self.distance_weight = float(distance_weight)
        self.match_mode = int(match_mode)
        self.x = float(x)
        self.y = float(y)
        self.z = float(z)

    @classmethod
    def loads(cls, line: str) -> "TemplateAtom":
        """Parse one 66-column template ``ATOM`` line; ``ValueError`` on anything else
        (re-raised by the loader at ``enzymm/template.py:1549-1552``)."""
        line = line.rstrip("\r\n")
        if not line.startswith("ATOM") or len(line) < 54:
            raise ValueError(f"not a template ATOM line: {line!r}")
        try:
            mode = int(line[6:11])
            resnum = int(line[22:26])
            x = float(line[30:38])
            y = float(line[38:46])
            z = float(line[46:54])
        except ValueError as exc:
            raise ValueError(f"malformed template ATOM line: {line!r}") from exc
        name = line[12:16].strip()
        resname = line[17:20].strip()
        if not name or not resname:
            raise ValueError(f"template ATOM line lacks atom or residue name: {line!r}")
        names = [resname]
        for letter in line[54:60].strip():
            three = ONE_TO_THREE.get(letter)
            if three is None:
                raise ValueError(f"unknown alternate residue code {letter!r} in template line {line!r}")
            if three not in names:
                names.append(three)
        weight_txt = line[60:66].strip()
        try:
            weight = float(weight_txt) if weight_txt else 0.0
        except ValueError as exc:
            raise ValueError(f"malformed distance weight in template line {line!r}") from exc
        return cls(chain_id=line[20:22].strip(), residue_number=resnum, residue_names=names,
                   atom_names=[name], distance_weight=weight, match_mode=mode, x=x, y=y, z=z)

    def _key(self) -> Tuple:
        return (self.chain_id, self.residue_number, self.residue_names, self.atom_names,
                self.distance_weight, self.match_mode, self.x, self.y, self.z)

    def typing_key(self) -> Tuple:
        """What decides which query atoms this atom may bind (SURVEY §8c rules 2-3)."""
        return (self.match_mode, self.residue_names, self.atom_names)

    def copy(self) -> "TemplateAtom":
        return TemplateAtom(chain_id=self.chain_id, residue_number=self.residue_number,
                            residue_names=self.residue_names, atom_names=self.atom_names,
                            distance_weight=self.distance_weight, match_mode=self.match_mode,
                            x=self.x, y=self.y, z=self.z)

    def __eq__(self, other):
        if not isinstance(other, TemplateAtom):
            return NotImplemented
        return self._key() == other._key()

    def __ne__(self, other):
        if not isinstance(other, TemplateAtom):
            return NotImplemented
        return self._key() != other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return (f"TemplateAtom(chain_id={self.chain_id!r}, residue_number={self.residue_number}, "
                f"residue_names={list(self.residue_names)}, atom_names={list(self.atom_names)}, "
                f"match_mode={self.match_mode}, x={self.x}, y={self.y}, z={self.z})")


class JessTemplate:
    """An ordered list of template atoms with an id (mirror of ``pyjess.Template``).

    Subclassable: EnzyMM's ``Template`` derives from it and adds metadata
    (``enzymm/template.py:455-551``); ``Hit.template`` hands back the very object given to
    ``Jess`` (``jess_run.py:162,203``)."""

    def __init__(self, atoms: Iterable[TemplateAtom] = (), id: Optional[str] = None):
        self._atoms: Tuple[TemplateAtom, ...] = tuple(atoms)
        for a in self._atoms:
            if not isinstance(a, TemplateAtom):
                raise TypeError(f"expected TemplateAtom, got {type(a).__name__}")
        self.id = id

    @property
    def dimension(self) -> int:
        """Number of distinct template residues (pinned 5 / 6 / 4 by ``tests/golden/results.tsv``
        and ``tests/test_template.py:358,388`` of the reference)."""
        return len({(a.chain_id, a.residue_number) for a in self._atoms})

    def __len__(self) -> int:
        return len(self._atoms)

    def __iter__(self) -> Iterator[TemplateAtom]:
        return iter(self._atoms)

    def __getitem__(self, i):
        return self._atoms[i]

    def copy(self) -> "JessTemplate":
        return JessTemplate(self._atoms, id=self.id)

    def __eq__(self, other):
        if not isinstance(other, JessTemplate):
            return NotImplemented
        return self.id == other.id and self._atoms == other._atoms

    def __hash__(self):
        return hash((self.id, self._atoms))
