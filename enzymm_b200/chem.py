"""Standard amino-acid heavy-atom topology and background composition.

Used for two things only: (1) to pre-register the typing classes of ordinary protein atoms when a
template library is compiled (so the device typing matrix never has to grow for protein input)
and to estimate how selective each template atom is when the search plan is ordered;
(2) by the synthetic structure generator (``enzymm_b200/synth.py``).
"""
from __future__ import annotations

from typing import Dict, Tuple

BACKBONE = ("N", "CA", "C", "O")

SIDE_CHAINS: Dict[str, Tuple[str, ...]] = {
    "ALA": ("CB",),
    "ARG": ("CB", "CG", "CD", "NE", "CZ", "NH1", "NH2"),
    "ASN": ("CB", "CG", "OD1", "ND2"),
    "ASP": ("CB", "CG", "OD1", "OD2"),
    "CYS": ("CB", "SG"),
    "GLN": ("CB", "CG", "CD", "OE1", "NE2"),
    "GLU": ("CB", "CG", "CD", "OE1", "OE2"),
    "GLY": (),
    "HIS": ("CB", "CG", "ND1", "CD2", "CE1", "NE2"),
    "ILE": ("CB", "CG1", "CG2", "CD1"),
    "LEU": ("CB", "CG", "CD1", "CD2"),
    "LYS": ("CB", "CG", "CD", "CE", "NZ"),
    "MET": ("CB", "CG", "SD", "CE"),
    "PHE": ("CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ"),
    "PRO": ("CB", "CG", "CD"),
    "SER": ("CB", "OG"),
    "THR": ("CB", "OG1", "CG2"),
    "TRP": ("CB", "CG", "CD1", "CD2", "NE1", "CE2", "CE3", "CZ2", "CZ3", "CH2"),
    "TYR": ("CB", "CG", "CD1", "CD2", "CE1", "CE2", "CZ", "OH"),
    "VAL": ("CB", "CG1", "CG2"),
}

RESIDUE_ATOMS: Dict[str, Tuple[str, ...]] = {r: BACKBONE + sc for r, sc in SIDE_CHAINS.items()}

# UniProtKB/Swiss-Prot amino-acid composition (percent)
BACKGROUND_PERCENT: Dict[str, float] = {
    "ALA": 8.25, "ARG": 5.53, "ASN": 4.06, "ASP": 5.45, "CYS": 1.37, "GLN": 3.93, "GLU": 6.75,
    "GLY": 7.07, "HIS": 2.27, "ILE": 5.96, "LEU": 9.66, "LYS": 5.84, "MET": 2.42, "PHE": 3.86,
    "PRO": 4.70, "SER": 6.56, "THR": 5.34, "TRP": 1.08, "TYR": 2.92, "VAL": 6.87,
}

RESIDUE_ORDER = tuple(sorted(RESIDUE_ATOMS))


def element_of(atom_name: str) -> str:
    """Element symbol of a standard protein atom name."""
    return atom_name[0]
