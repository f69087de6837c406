"""Deterministic synthetic query structures with planted M-CSA motifs.

AlphaFold-scale stand-ins for the throughput configs of BASELINE.json: a self-avoiding,
compactness-biased CA random walk (3.8 A steps) dressed with heavy atoms from a frozen
per-residue conformation table (real PDB atom names, ~7.8 atoms/residue), residue types drawn
from the UniProt background, pLDDT-like B-factors, and 0-3 templates of the searched library
planted rigidly with isotropic Gaussian noise into residues of an allowed type.

Determinism: structures are produced in fixed chunks of ``CHUNK`` structures, chunk ``c`` drawing
from ``Philox(key=[seed, c])`` -- so structure ``i`` is the same whoever generates it, and ranks can
shard by chunk.  Coordinates are rounded to 3 decimals (as PDB text stores them), so the packed
arrays and the PDB text describe bit-identical doubles.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .chem import BACKGROUND_PERCENT, RESIDUE_ATOMS, RESIDUE_ORDER
from .engine import PackedBatch
from .library import CompiledLibrary, chain_code
from .structures import Molecule, _COLUMNS

__all__ = ["SynthConfig", "SynthChunk", "generate_chunk", "generate_batch", "CHUNK", "SEED"]

SEED = 20230210
CHUNK = 256
_GEOM = json.loads((Path(__file__).resolve().parent / "data" / "residue_geometry.json").read_text())

# flat table of atom kinds: kind id -> (residue type index, atom name), grouped by residue type
_KIND_RES: List[int] = []
_KIND_NAME: List[str] = []
_KIND_LOCAL: List[Tuple[float, float, float]] = []
_RES_FIRST: List[int] = []
_RES_COUNT: List[int] = []
for _ri, _res in enumerate(RESIDUE_ORDER):
    _RES_FIRST.append(len(_KIND_RES))
    _RES_COUNT.append(len(RESIDUE_ATOMS[_res]))
    for _name in RESIDUE_ATOMS[_res]:
        _KIND_RES.append(_ri)
        _KIND_NAME.append(_name)
        _KIND_LOCAL.append(tuple(_GEOM[_res][_name]))
_KIND_LOCAL_ARR = np.asarray(_KIND_LOCAL, dtype=np.float64)
_RES_FIRST_ARR = np.asarray(_RES_FIRST, dtype=np.int64)
_RES_COUNT_ARR = np.asarray(_RES_COUNT, dtype=np.int64)
_FREQ = np.asarray([BACKGROUND_PERCENT[r] for r in RESIDUE_ORDER], dtype=np.float64)
_FREQ /= _FREQ.sum()
_RES_INDEX = {r: i for i, r in enumerate(RESIDUE_ORDER)}


@dataclass(frozen=True)
class SynthConfig:
    n_residues: int = 400          # per chain
    n_chains: int = 1
    seed: int = SEED
    max_motifs: int = 3            # 0..max_motifs planted templates per structure
    noise_sigma: float = 0.3       # Angstrom, per coordinate, on planted atoms
    compactness: float = 0.0007    # restoring-force weight of the walk (tuned to protein-like density)


@dataclass
class SynthChunk:
    """SoA columns of a run of synthetic structures."""

    atom_off: np.ndarray       # int64 [S+1]
    xyz: np.ndarray            # float64 [N,3], 3-decimal values
    kind: np.ndarray           # int16 [N] index into the atom-kind table
    residue: np.ndarray        # int32 [N] residue ordinal within the structure
    resnum: np.ndarray         # int32 [N] PDB residue number (1-based, per chain)
    chain: np.ndarray          # uint16 [N] chain code
    bfactor: np.ndarray        # float32 [N]
    planted: List[Tuple[int, int]]   # (structure index within chunk list, template index)
    first_index: int = 0       # global index of structure 0

    @property
    def n_structures(self) -> int:
        return len(self.atom_off) - 1

    @property
    def n_atoms(self) -> int:
        return int(self.atom_off[-1])

    def to_packed(self, library: CompiledLibrary, with_chain: bool = False) -> PackedBatch:
        return PackedBatch(self.atom_off, self.xyz, kind_classes(library)[self.kind], self.residue, self.bfactor,
                           self.chain if with_chain else None, None)

    def to_molecule(self, i: int) -> Molecule:
        lo, hi = int(self.atom_off[i]), int(self.atom_off[i + 1])
        kind = self.kind[lo:hi]
        names = np.asarray(_KIND_NAME)[kind]
        resn = np.asarray(RESIDUE_ORDER)[np.asarray(_KIND_RES)[kind]]
        chains = np.asarray([chr(c & 0xFF) for c in self.chain[lo:hi]])
        n = hi - lo
        cols = {
            "serial": np.arange(1, n + 1, dtype=np.int32), "name": names.astype("U4"),
            "altloc": np.full(n, " ", dtype="U1"), "residue_name": resn.astype("U4"),
            "chain_id": chains.astype("U2"), "residue_number": self.resnum[lo:hi].astype(np.int32),
            "insertion_code": np.full(n, " ", dtype="U1"), "occupancy": np.ones(n, dtype=np.float64),
            "temperature_factor": self.bfactor[lo:hi].astype(np.float64), "segment": np.full(n, "", dtype="U4"),
            "element": np.asarray([s[0] for s in names]).astype("U2"), "charge": np.zeros(n, dtype=np.int8),
        }
        assert set(cols) == {k for k, _ in _COLUMNS}
        return Molecule._from_columns(cols, self.xyz[lo:hi].copy(), f"synth_{self.first_index + i:07d}")

    def to_pdb(self, i: int) -> str:
        """PDB text of structure ``i`` (so a real reference run could consume the same input)."""
        mol = self.to_molecule(i)
        lines = [f"HEADER    SYNTHETIC STRUCTURE {mol.id}\n"]
        for a in mol:
            name = f" {a.name:<3s}" if len(a.name) < 4 else a.name
            lines.append(f"ATOM  {a.serial:>5} {name}{a.altloc}{a.residue_name:>3}{a.chain_id:>2}{a.residue_number:>4}"
                         f"{a.insertion_code}   {a.x:>8.3f}{a.y:>8.3f}{a.z:>8.3f}{a.occupancy:>6.2f}"
                         f"{a.temperature_factor:>6.2f}          {a.element:>2s}\n")
        lines.append("END\n")
        return "".join(lines)


def kind_classes(library: CompiledLibrary) -> np.ndarray:
    """Typing class (of ``library``) of every synthetic atom kind: ``klass = kind_classes(lib)[chunk.kind]``."""
    return np.asarray([library.class_of(RESIDUE_ORDER[r], n) for r, n in zip(_KIND_RES, _KIND_NAME)], dtype=np.uint16)


def _unit(v: np.ndarray) -> np.ndarray:
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    return v / np.where(n == 0, 1.0, n)


def _walk(rng: np.random.Generator, S: int, L: int, lam: float, start: np.ndarray) -> np.ndarray:
    """Vectorised self-avoiding CA walk for S chains of L residues.

    Each step draws K candidate directions and keeps the one minimising
    ``lam * r^2`` (compactness) + occupancy (a 1.5 A voxel grid on which every placed CA stamps a
    4.5 A sphere) + a bend penalty (CA(i-2)..CA(i) >= 5 A).  Measured on 256 x 400 residues:
    radius of gyration 21.2 +- 1.2 A and ~0.014 heavy atoms / A^3 of bounding box, as the fixtures.
    """
    K, G, VOX, RAD = 10, 80, 1.5, 4.5
    r = int(np.ceil(RAD / VOX))
    ax = np.arange(-r, r + 1)
    da, db, dc = np.meshgrid(ax, ax, ax, indexing="ij")
    inside = (da * da + db * db + dc * dc) * VOX * VOX <= RAD * RAD
    stamp = ((da[inside] * G + db[inside]) * G + dc[inside]).astype(np.int64)
    P = np.zeros((S, L, 3), dtype=np.float64)
    P[:, 0] = start
    occ = np.zeros((S, G * G * G), dtype=np.int8)
    rows = np.arange(S)

    def voxel(p):
        rel = p - (start[:, None, :] if p.ndim == 3 else start)
        idx = np.clip(np.floor(rel / VOX).astype(np.int64) + G // 2, r, G - 1 - r)
        return (idx[..., 0] * G + idx[..., 1]) * G + idx[..., 2]

    for i in range(1, L):
        u = _unit(rng.standard_normal((S, K, 3)))
        cand = P[:, i - 1, None, :] + 3.8 * u
        rel = cand - start[:, None, :]
        energy = lam * np.einsum("skc,skc->sk", rel, rel)
        energy += 30.0 * occ[rows[:, None], voxel(cand)]
        if i >= 2:
            back = cand - P[:, i - 2, None, :]
            energy += np.where(np.einsum("skc,skc->sk", back, back) < 25.0, 50.0, 0.0)
        energy += 0.5 * rng.random((S, K))
        P[:, i] = cand[rows, np.argmin(energy, axis=1)]
        if i >= 2:   # stamp with a lag of two so a chain does not block its own next steps
            cells = voxel(P[:, i - 2])[:, None] + stamp[None, :]
            occ[rows[:, None], cells] = np.minimum(occ[rows[:, None], cells] + 1, 100)
    return P


def _frames(P: np.ndarray) -> np.ndarray:
    """Per-residue rotation matrices [S, L, 3, 3] (columns = local x, y, z axes) from the CA trace."""
    nxt = np.concatenate([P[:, 1:], 2 * P[:, -1:] - P[:, -2:-1]], axis=1)
    prv = np.concatenate([2 * P[:, :1] - P[:, 1:2], P[:, :-1]], axis=1)
    x = _unit(nxt - P)
    v = prv - P
    y = v - np.einsum("slc,slc->sl", v, x)[..., None] * x
    bad = np.linalg.norm(y, axis=-1) < 1e-6
    if bad.any():
        alt = np.cross(x, np.array([0.0, 0.0, 1.0]))
        y = np.where(bad[..., None], alt, y)
    y = _unit(y)
    z = np.cross(x, y)
    return np.stack([x, y, z], axis=-1)


def _kabsch_fit(src: np.ndarray, dst: np.ndarray):
    """Rigid transform (R, t) with R @ src_i + t ~= dst_i (proper rotation)."""
    cs, cd = src.mean(axis=0), dst.mean(axis=0)
    H = (src - cs).T @ (dst - cd)
    U, _, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(Vt.T @ U.T))
    R = Vt.T @ np.diag([1.0, 1.0, d]) @ U.T
    return R, cd - R @ cs


def _random_rotation(rng: np.random.Generator) -> np.ndarray:
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _host_type(atoms) -> Optional[str]:
    """A standard residue type that can host this template residue's three atoms."""
    wanted = [a.atom_names[0] for a in atoms]
    for res in atoms[0].residue_names:
        if res in RESIDUE_ATOMS and all(n in RESIDUE_ATOMS[res] for n in wanted):
            return res
    return None


def generate_chunk(chunk_index: int, cfg: SynthConfig = SynthConfig(),
                   templates: Optional[Sequence] = None, count: int = CHUNK) -> SynthChunk:
    """Structures ``chunk_index*CHUNK .. +count`` (``count <= CHUNK``)."""
    rng = np.random.Generator(np.random.Philox(key=[cfg.seed, chunk_index]))
    S, L, C = CHUNK, cfg.n_residues, cfg.n_chains
    LT = L * C
    # chain origins on a tetrahedron-like arrangement ~ one globule diameter apart
    offsets = np.array([[0, 0, 0], [38, 0, 0], [19, 33, 0], [19, 11, 31]], dtype=np.float64)
    if C > len(offsets):
        raise ValueError("at most 4 chains")
    P = np.concatenate([_walk(rng, S, L, cfg.compactness, np.tile(offsets[c], (S, 1))) for c in range(C)], axis=1)
    restype = rng.choice(len(RESIDUE_ORDER), size=(S, LT), p=_FREQ)
    plddt = np.round(30.0 + 69.0 * rng.beta(5.0, 2.0, size=(S, LT)), 2)
    frames = _frames(P.reshape(S * C, L, 3)).reshape(S, LT, 3, 3)

    # ---- plant motifs: decide hosts and residue types before atoms are laid out ------------------
    overrides: Dict[Tuple[int, int], Tuple[np.ndarray, np.ndarray, List[Tuple[str, np.ndarray]]]] = {}
    planted: List[Tuple[int, int]] = []
    n_motifs = rng.integers(0, cfg.max_motifs + 1, size=S) if templates else np.zeros(S, dtype=np.int64)
    for s in range(S):
        for _ in range(int(n_motifs[s])):
            ti = int(rng.integers(len(templates)))
            rot = _random_rotation(rng)
            anchor = int(rng.integers(LT))
            noise_seed = rng.standard_normal((32, 3))
            if s >= count:
                continue   # random stream stays aligned whatever `count` is
            tpl = templates[ti]
            atoms = list(tpl)
            txyz = np.array([(a.x, a.y, a.z) for a in atoms])
            target = (txyz - txyz.mean(axis=0)) @ rot.T + P[s, anchor]
            groups: Dict[tuple, List[int]] = {}
            for i, a in enumerate(atoms):
                groups.setdefault((a.chain_id, a.residue_number), []).append(i)
            taken = {r for (ss, r) in overrides if ss == s}
            plan = []
            ok = True
            for idxs in groups.values():
                res = _host_type([atoms[i] for i in idxs]) if len(idxs) == 3 else None
                if res is None:
                    ok = False
                    break
                centre = target[idxs].mean(axis=0)
                order = np.argsort(np.einsum("lc,lc->l", P[s] - centre, P[s] - centre))
                host = next((int(r) for r in order if int(r) not in taken), None)
                if host is None:
                    ok = False
                    break
                taken.add(host)
                plan.append((host, res, idxs))
            if not ok:
                continue
            for host, res, idxs in plan:
                names = [atoms[i].atom_names[0] for i in idxs]
                local = np.array([_GEOM[res][n] for n in names])
                goal = target[idxs] + cfg.noise_sigma * noise_seed[idxs]
                R, t = _kabsch_fit(local, goal)
                restype[s, host] = _RES_INDEX[res]
                overrides[(s, host)] = (R, t, list(zip(names, goal)))
            planted.append((s, ti))

    # ---- lay atoms out -------------------------------------------------------------------------------
    S_out = count
    rt = restype[:S_out].reshape(-1)
    counts = _RES_COUNT_ARR[rt]
    res_of_atom = np.repeat(np.arange(S_out * LT), counts)
    first = np.cumsum(counts) - counts
    within = np.arange(len(res_of_atom)) - np.repeat(first, counts)
    kind = (_RES_FIRST_ARR[rt][res_of_atom] + within).astype(np.int16)
    local = _KIND_LOCAL_ARR[kind]
    F = frames[:S_out].reshape(-1, 3, 3)[res_of_atom]
    xyz = P[:S_out].reshape(-1, 3)[res_of_atom] + np.einsum("nij,nj->ni", F, local)
    for (s, host), (R, t, exact) in overrides.items():
        if s >= S_out:
            continue
        r = s * LT + host
        lo, hi = int(first[r]), int(first[r] + counts[r])
        xyz[lo:hi] = _KIND_LOCAL_ARR[kind[lo:hi]] @ R.T + t
        names = [_KIND_NAME[k] for k in kind[lo:hi]]
        for name, pos in exact:
            xyz[lo + names.index(name)] = pos
    xyz = np.round(xyz, 3)
    atoms_per_structure = np.add.reduceat(counts, np.arange(0, S_out * LT, LT))
    atom_off = np.zeros(S_out + 1, dtype=np.int64)
    np.cumsum(atoms_per_structure, out=atom_off[1:])
    res_in_structure = (res_of_atom % LT).astype(np.int32)
    chain_idx = res_in_structure // L
    codes = np.asarray([chain_code(chr(ord("A") + c)) for c in range(C)], dtype=np.uint16)
    return SynthChunk(atom_off=atom_off, xyz=xyz, kind=kind, residue=res_in_structure,
                      resnum=(res_in_structure % L + 1).astype(np.int32), chain=codes[chain_idx],
                      bfactor=plddt[:S_out].reshape(-1)[res_of_atom].astype(np.float32),
                      planted=[p for p in planted if p[0] < S_out], first_index=chunk_index * CHUNK)


def generate_batch(first: int, count: int, cfg: SynthConfig = SynthConfig(),
                   templates: Optional[Sequence] = None) -> SynthChunk:
    """Structures ``first .. first+count`` (``first`` must be a multiple of ``CHUNK``)."""
    if first % CHUNK:
        raise ValueError(f"first must be a multiple of {CHUNK}")
    parts: List[SynthChunk] = []
    done = 0
    while done < count:
        n = min(CHUNK, count - done)
        parts.append(generate_chunk((first + done) // CHUNK, cfg, templates, n))
        done += n
    if len(parts) == 1:
        return parts[0]
    atom_off = [np.zeros(1, dtype=np.int64)]
    planted: List[Tuple[int, int]] = []
    base_s, base_a = 0, 0
    for p in parts:
        atom_off.append(p.atom_off[1:] + base_a)
        planted.extend((s + base_s, t) for s, t in p.planted)
        base_s += p.n_structures
        base_a += p.n_atoms
    cat = lambda name: np.concatenate([getattr(p, name) for p in parts])
    return SynthChunk(np.concatenate(atom_off), cat("xyz"), cat("kind"), cat("residue"), cat("resnum"),
                      cat("chain"), cat("bfactor"), planted, first_index=first)
