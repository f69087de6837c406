"""Freeze one heavy-atom conformation per standard residue type into
``enzymm_b200/data/residue_geometry.json`` (used by the synthetic structure generator).

Source: the reference's own fixture ``tests/golden/1AMY.pdb`` (a crystal structure), first complete
instance of each residue type.  Coordinates are expressed in the residue's backbone frame:
origin CA, x along CA->C, y = component of CA->N orthogonal to x, z = x cross y.

Run once:  python tools/make_residue_geometry.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from enzymm_b200.chem import RESIDUE_ATOMS  # noqa: E402
from enzymm_b200.structures import Molecule  # noqa: E402


def main():
    mol = Molecule.load(ROOT / "tests/golden/1AMY.pdb")
    names, resn, resi = mol.column("name"), mol.column("residue_name"), mol.column("residue_number")
    out = {}
    for res, atoms in RESIDUE_ATOMS.items():
        for num in np.unique(resi[resn == res]):
            sel = np.nonzero((resi == num) & (resn == res))[0]
            have = {str(names[i]): mol.xyz[i] for i in sel}
            if all(a in have for a in atoms):
                ca, c, n = have["CA"], have["C"], have["N"]
                x = (c - ca) / np.linalg.norm(c - ca)
                y = (n - ca) - np.dot(n - ca, x) * x
                y /= np.linalg.norm(y)
                z = np.cross(x, y)
                frame = np.stack([x, y, z])          # rows = axes
                out[res] = {a: [round(float(v), 3) for v in frame @ (have[a] - ca)] for a in atoms}
                break
        else:
            raise SystemExit(f"no complete {res} in the fixture")
    dst = ROOT / "enzymm_b200/data/residue_geometry.json"
    dst.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")
    print(f"wrote {dst} ({len(out)} residue types)")


if __name__ == "__main__":
    main()
