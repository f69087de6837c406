"""Molecules -> ``PackedBatch``: the host half of north_star subsystem (1), batched upload.

Turns ``Molecule`` objects (the ``pyjess.Molecule`` stand-in) into the SoA columns of
``emm_batch``: float64 coordinates exactly as parsed, typing class per atom, residue ordinal per
atom (a residue = one (chain_id, residue_number), SURVEY.md 8c rule 4), B-factor, chain code.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .engine import PackedBatch
from .library import CompiledLibrary
from .structures import Molecule

__all__ = ["residue_ordinals", "pack_molecules", "chain_codes"]


def chain_codes(col: np.ndarray) -> np.ndarray:
    """U2 chain ids -> uint16 codes (byte0 | byte1 << 8), matching ``library.chain_code``."""
    if len(col) == 0:
        return np.zeros(0, dtype=np.uint16)
    raw = np.char.encode(col.astype("U2"), "ascii", "replace").astype("S2")
    b = np.frombuffer(raw.tobytes(), dtype=np.uint8).reshape(-1, 2).astype(np.uint16)
    return b[:, 0] | (b[:, 1] << 8)


def residue_ordinals(chain: np.ndarray, resnum: np.ndarray):
    """Residue ordinal per atom, numbered by first appearance.

    Returns ``(ordinal, order)``: ``order`` is None when every residue's atoms are already
    contiguous (every PDB file in practice); otherwise it is the stable permutation that makes
    them contiguous, and ``ordinal`` refers to the permuted atoms.
    """
    n = len(resnum)
    if n == 0:
        return np.zeros(0, dtype=np.int32), None
    key = (chain.astype(np.int64) << 32) | (resnum.astype(np.int64) & 0xFFFFFFFF)
    change = np.empty(n, dtype=bool)
    change[0] = True
    np.not_equal(key[1:], key[:-1], out=change[1:])
    runs = np.cumsum(change) - 1
    n_runs = int(runs[-1]) + 1
    uniq, first_idx, inverse = np.unique(key, return_index=True, return_inverse=True)
    if len(uniq) == n_runs:
        return runs.astype(np.int32), None
    # some residue is split over several runs: number residues by first appearance and sort
    rank_of_uniq = np.empty(len(uniq), dtype=np.int64)
    rank_of_uniq[np.argsort(first_idx, kind="stable")] = np.arange(len(uniq))
    ordinal = rank_of_uniq[inverse]
    order = np.argsort(ordinal, kind="stable")
    return ordinal[order].astype(np.int32), order.astype(np.int32)


def pack_molecules(molecules: Sequence[Molecule], library: CompiledLibrary,
                   with_chain: bool = True) -> PackedBatch:
    """Concatenate molecules into one ``PackedBatch`` (classes come from ``library.classify``)."""
    sizes = [len(m) for m in molecules]
    atom_off = np.zeros(len(molecules) + 1, dtype=np.int64)
    np.cumsum(sizes, out=atom_off[1:])
    total = int(atom_off[-1])
    xyz = np.empty((total, 3), dtype=np.float64)
    klass = np.empty(total, dtype=np.uint16)
    residue = np.empty(total, dtype=np.int32)
    bfactor = np.empty(total, dtype=np.float32)
    chain = np.empty(total, dtype=np.uint16)
    atom_id: Optional[np.ndarray] = None
    for i, m in enumerate(molecules):
        lo, hi = int(atom_off[i]), int(atom_off[i + 1])
        if hi == lo:
            continue
        codes = chain_codes(m.column("chain_id"))
        ordinal, order = residue_ordinals(codes, m.column("residue_number"))
        kl = library.classify(m.column("residue_name"), m.column("name"))
        bf = m.column("temperature_factor").astype(np.float32)
        coords = m.xyz
        if order is not None:
            if atom_id is None:
                atom_id = np.empty(total, dtype=np.int32)
                for j in range(i):
                    atom_id[atom_off[j]:atom_off[j + 1]] = np.arange(sizes[j], dtype=np.int32)
            atom_id[lo:hi] = order
            coords, kl, bf, codes = coords[order], kl[order], bf[order], codes[order]
        elif atom_id is not None:
            atom_id[lo:hi] = np.arange(hi - lo, dtype=np.int32)
        xyz[lo:hi] = coords
        klass[lo:hi] = kl
        residue[lo:hi] = ordinal
        bfactor[lo:hi] = bf
        chain[lo:hi] = codes
    return PackedBatch(atom_off, xyz, klass, residue, bfactor, chain if with_chain else None, atom_id)
