"""CPU oracle for the geometric matching hot path -- TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``enzymm_b200`` never does.

``jess_oracle.c`` restates what ``pyjess.Jess(templates).query(...)`` computes (call site
``enzymm/jess_run.py:800-811``; the implementation lives in the un-vendored dependency
``pyjess ~=0.5.0``, reference ``pyproject.toml:30``).  This module is its ctypes binding plus a
plain-Python restatement of EnzyMM's own post-processing:

  * ``orientation``        -- ``enzymm/jess_run.py:348-373, 425-478`` + ``template.py:157-181``
  * ``predicted_correct``  -- ``enzymm/jess_run.py:298-346``

Parity status: PINNED against every golden vector the reference's tests hold for this path
(``tests/test_oracle_golden.py``); ``match_mode 1``, dynamic distances, ``ignore_chain=False``
and ``max_candidates`` truncation order are UNPINNED (no reference vector exists).
"""
from __future__ import annotations

import ctypes
import json
import math
import os
import subprocess
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libjess_oracle.so"
MAX_ATOMS = 32


class _Result(ctypes.Structure):
    _fields_ = [
        ("found", ctypes.c_int32), ("overflow", ctypes.c_int32), ("rmsd", ctypes.c_double),
        ("atoms", ctypes.c_int32 * MAX_ATOMS), ("rot", ctypes.c_double * 9),
        ("qbar", ctypes.c_double * 3), ("tbar", ctypes.c_double * 3),
        ("n_complete", ctypes.c_int64), ("n_accepted", ctypes.c_int64),
        ("nodes", ctypes.c_int64), ("dist_evals", ctypes.c_int64),
    ]


_RESULT_DTYPE = np.dtype([
    ("found", np.int32), ("overflow", np.int32), ("rmsd", np.float64),
    ("atoms", np.int32, (MAX_ATOMS,)), ("rot", np.float64, (9,)), ("qbar", np.float64, (3,)),
    ("tbar", np.float64, (3,)), ("n_complete", np.int64), ("n_accepted", np.int64),
    ("nodes", np.int64), ("dist_evals", np.int64),
])


def build(force: bool = False) -> Path:
    """Compile ``jess_oracle.c`` with the committed Makefile (gcc, no external libraries)."""
    src = _HERE / "jess_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "libjess_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(str(_LIB_PATH))
        assert lib.jo_result_size() == ctypes.sizeof(_Result) == _RESULT_DTYPE.itemsize
        assert lib.jo_max_atoms() == MAX_ATOMS
        lib.jo_batch_query.restype = ctypes.c_int
        lib.jo_kabsch.restype = ctypes.c_double
        _lib = lib
        reading = os.environ.get("EMM_ORACLE_MODE1")      # exposure studies only; the default is "N_or_O"
        if reading:
            lib.jo_set_mode1_reading(ctypes.c_int(MODE1_READINGS.index(reading)))
    return _lib


def _fixed(strings: Sequence[str], width: int) -> np.ndarray:
    out = np.zeros((len(strings), width), dtype=np.uint8)
    for i, s in enumerate(strings):
        b = s.encode("ascii", "replace")[:width]
        out[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    return out


def _fixed_col(col: np.ndarray, width: int) -> np.ndarray:
    """NumPy unicode column -> [n, width] NUL padded ASCII bytes."""
    if len(col) == 0:
        return np.zeros((0, width), dtype=np.uint8)
    as_bytes = np.char.encode(col.astype(f"U{width}"), "ascii", "replace").astype(f"S{width}")
    return np.frombuffer(as_bytes.tobytes(), dtype=np.uint8).reshape(len(col), width).copy()


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class OracleTemplates:
    """Flattened template batch (strings kept as strings: typing is done in C on the names)."""

    def __init__(self, templates: Sequence):
        self.templates = list(templates)
        offs = [0]
        xyz, mode, chain, resnum, weight = [], [], [], [], []
        an_off, rn_off, an_pool, rn_pool = [0], [0], [], []
        keys: Dict[tuple, int] = {}
        tkey, key_rep = [], []
        for t in self.templates:
            atoms = list(t)
            if not 0 < len(atoms) <= MAX_ATOMS:
                raise ValueError(f"template with {len(atoms)} atoms (oracle limit {MAX_ATOMS})")
            for a in atoms:
                k = (a.match_mode, tuple(a.residue_names), tuple(a.atom_names))
                if k not in keys:
                    keys[k] = len(keys)
                    key_rep.append(len(xyz))
                tkey.append(keys[k])
                xyz.append((a.x, a.y, a.z))
                mode.append(a.match_mode)
                chain.append(a.chain_id)
                resnum.append(a.residue_number)
                weight.append(a.distance_weight)
                an_pool.extend(a.atom_names)
                rn_pool.extend(a.residue_names)
                an_off.append(len(an_pool))
                rn_off.append(len(rn_pool))
            offs.append(len(xyz))
        self.tpl_off = np.asarray(offs, dtype=np.int32)
        self.xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
        self.mode = np.asarray(mode, dtype=np.int32)
        self.chain = _fixed(chain, 2)
        self.resnum = np.asarray(resnum, dtype=np.int32)
        self.weight = np.asarray(weight, dtype=np.float64)
        self.an_off = np.asarray(an_off, dtype=np.int32)
        self.rn_off = np.asarray(rn_off, dtype=np.int32)
        self.an_pool = _fixed(an_pool, 4)
        self.rn_pool = _fixed(rn_pool, 4)
        self.tkey = np.asarray(tkey, dtype=np.int32)
        self.key_rep = np.asarray(key_rep, dtype=np.int32)

    def __len__(self):
        return len(self.templates)


@dataclass
class OracleHit:
    molecule_index: int
    template_index: int
    rmsd: float
    atoms: List[int]
    rot: np.ndarray          # 3x3, query frame -> template frame
    qbar: np.ndarray
    tbar: np.ndarray
    n_complete: int
    n_accepted: int
    overflow: bool

    def transform(self, xyz: np.ndarray) -> np.ndarray:
        """q' = R (q - qbar) + tbar  (SURVEY 8c rule 9)."""
        return (np.asarray(xyz, dtype=np.float64) - self.qbar) @ self.rot.T + self.tbar


MODE1_READINGS = ("N_or_O", "same_element", "exact_name", "N_O_or_S")


def set_mode1_reading(reading: str = "N_or_O") -> None:
    """Reading of ``match_mode 1`` (unpinned upstream): the default is "N_or_O"; the others
    exist to measure how much of the result depends on the choice (``tools/mode1_exposure.py``)."""
    _load().jo_set_mode1_reading(ctypes.c_int(MODE1_READINGS.index(reading)))


def _per_template(value, n: int) -> np.ndarray:
    arr = np.asarray(value, dtype=np.float64)
    if arr.ndim == 0:
        arr = np.full(n, float(arr), dtype=np.float64)
    if arr.shape != (n,):
        raise ValueError("per-template parameter has wrong length")
    return np.ascontiguousarray(arr)


def query_raw(molecules: Sequence, templates: OracleTemplates, rmsd_threshold, distance_cutoff,
              max_dynamic_distance, max_candidates: int = 10000, ignore_chain: bool = True,
              threads: int = 1) -> np.ndarray:
    """Run the oracle; returns the raw ``[n_mol, n_tpl]`` structured result array."""
    lib = _load()
    n_mol, n_tpl = len(molecules), len(templates)
    sizes = [len(m) for m in molecules]
    mol_off = np.zeros(n_mol + 1, dtype=np.int64)
    np.cumsum(sizes, out=mol_off[1:])
    total = int(mol_off[-1])
    xyz = np.concatenate([m.xyz for m in molecules]).astype(np.float64) if total else np.zeros((0, 3))
    name = np.concatenate([_fixed_col(m.column("name"), 4) for m in molecules]) if total else np.zeros((0, 4), np.uint8)
    resname = np.concatenate([_fixed_col(m.column("residue_name"), 4) for m in molecules]) if total else np.zeros((0, 4), np.uint8)
    chain = np.concatenate([_fixed_col(m.column("chain_id"), 2) for m in molecules]) if total else np.zeros((0, 2), np.uint8)
    resnum = np.concatenate([m.column("residue_number") for m in molecules]).astype(np.int32) if total else np.zeros(0, np.int32)
    xyz = np.ascontiguousarray(xyz)
    name, resname, chain = map(np.ascontiguousarray, (name, resname, chain))
    results = np.zeros((n_mol, n_tpl), dtype=_RESULT_DTYPE)
    rt = _per_template(rmsd_threshold, n_tpl)
    dc = _per_template(distance_cutoff, n_tpl)
    md = _per_template(max_dynamic_distance, n_tpl)
    t = templates
    rc = lib.jo_batch_query(
        ctypes.c_int(n_mol), _ptr(mol_off), _ptr(xyz), _ptr(name), _ptr(resname), _ptr(chain), _ptr(resnum),
        ctypes.c_int(n_tpl), _ptr(t.tpl_off), _ptr(t.xyz), _ptr(t.mode), _ptr(t.chain), _ptr(t.resnum),
        _ptr(t.weight), _ptr(t.an_off), _ptr(t.an_pool), _ptr(t.rn_off), _ptr(t.rn_pool),
        ctypes.c_int(len(t.key_rep)), _ptr(t.tkey), _ptr(t.key_rep),
        _ptr(rt), _ptr(dc), _ptr(md), ctypes.c_int64(int(max_candidates) if max_candidates else 0),
        ctypes.c_int(1 if ignore_chain else 0), ctypes.c_int(int(threads)), _ptr(results))
    if rc != 0:
        raise RuntimeError({-1: "oracle: out of memory", -2: "oracle: template too large or empty",
                            -3: "oracle: unknown match_mode"}.get(rc, f"oracle error {rc}"))
    return results


def query(molecules: Sequence, templates: OracleTemplates, rmsd_threshold, distance_cutoff,
          max_dynamic_distance, max_candidates: int = 10000, ignore_chain: bool = True,
          threads: int = 1) -> List[List[OracleHit]]:
    """Best hit per (molecule, template); per molecule a list ordered by template index."""
    raw = query_raw(molecules, templates, rmsd_threshold, distance_cutoff, max_dynamic_distance,
                    max_candidates, ignore_chain, threads)
    out: List[List[OracleHit]] = []
    for mi in range(raw.shape[0]):
        hits = []
        for ti in np.nonzero(raw[mi]["found"])[0]:
            r = raw[mi, ti]
            m = len(templates.templates[ti])
            hits.append(OracleHit(mi, int(ti), float(r["rmsd"]), [int(v) for v in r["atoms"][:m]],
                                  r["rot"].reshape(3, 3).copy(), r["qbar"].copy(), r["tbar"].copy(),
                                  int(r["n_complete"]), int(r["n_accepted"]), bool(r["overflow"])))
        out.append(hits)
    return out


def kabsch(template_xyz: np.ndarray, query_xyz: np.ndarray):
    """The oracle's superposition on its own: returns (rmsd, R, qbar, tbar)."""
    lib = _load()
    t = np.ascontiguousarray(template_xyz, dtype=np.float64)
    q = np.ascontiguousarray(query_xyz, dtype=np.float64)
    rot = np.zeros(9)
    qbar = np.zeros(3)
    tbar = np.zeros(3)
    rmsd = lib.jo_kabsch(ctypes.c_int(len(t)), _ptr(t), _ptr(q), _ptr(rot), _ptr(qbar), _ptr(tbar))
    return float(rmsd), rot.reshape(3, 3), qbar, tbar


# ---- EnzyMM post-processing restated in plain Python floats --------------------------------------

def _angle(u, v) -> float:
    """``Vec3.angle_to`` (enzymm/template.py:157-181)."""
    nu = math.sqrt(u[0] ** 2 + u[1] ** 2 + u[2] ** 2)
    nv = math.sqrt(v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
    a = u if nu == 0 else (u[0] / nu, u[1] / nu, u[2] / nu)
    b = v if nv == 0 else (v[0] / nv, v[1] / nv, v[2] / nv)
    dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
    if -1 <= dot <= 1:
        return math.acos(dot)
    if math.isclose(dot, 1, rel_tol=1e-5):
        return 0.0
    if math.isclose(dot, -1, rel_tol=1e-5):
        return math.pi
    raise ValueError("ArcCos is not defined outside [-1,1]")


def match_vectors(template, transformed_xyz: np.ndarray) -> List[tuple]:
    """Per-residue query orientation vectors in the template frame (jess_run.py:425-452).
    ``template.residues[i].orientation_vector_indices`` = (i, j) or (i, 9 = midpoint of others)."""
    vectors = []
    for ri, residue in enumerate(template.residues):
        a = [tuple(float(c) for c in transformed_xyz[3 * ri + k]) for k in range(3)]
        first, second = residue.orientation_vector_indices
        if second == 9:
            mid_atom = a[first]
            s1, s2 = [a[k] for k in range(3) if k != first]
            mid = ((s1[0] + s2[0]) / 2, (s1[1] + s2[1]) / 2, (s1[2] + s2[2]) / 2)
            vectors.append((mid[0] - mid_atom[0], mid[1] - mid_atom[1], mid[2] - mid_atom[2]))
        else:
            p, q = a[first], a[second]
            vectors.append((q[0] - p[0], q[1] - p[1], q[2] - p[2]))
    return vectors


def orientation(template, transformed_xyz: np.ndarray) -> float:
    """Mean angle between template and matched-residue orientation vectors (jess_run.py:461-478)."""
    mv = match_vectors(template, transformed_xyz)
    angles = []
    for residue, v in zip(template.residues, mv):
        tv = residue.orientation_vector
        angles.append(_angle((tv.x, tv.y, tv.z), v))
    return sum(angles) / len(angles)


_LR_MODELS: Optional[dict] = None


def lr_models(path: Optional[os.PathLike] = None) -> dict:
    """{size str: {distance str: [(coef0, coef1, intercept, threshold) x5]}} (jess_run.py:499-520)."""
    global _LR_MODELS
    if _LR_MODELS is None or path is not None:
        p = Path(path) if path else _HERE.parent / "enzymm_b200" / "data" / "logistic_regression_models.json"
        raw = json.loads(p.read_text())
        models = {}
        for size, by_dist in raw["match_size"].items():
            models[size] = {
                d: [(m["coef"][0], m["coef"][1], m["intercept"], m["threshold"]) for m in v["model_list"]]
                for d, v in by_dist["pairwise_distance"].items()
            }
        if path is not None:
            return models
        _LR_MODELS = models
    return _LR_MODELS


def predicted_correct(effective_size: int, pairwise_distance: float, rmsd: float, orient: float) -> bool:
    """Majority vote of the logistic models (jess_run.py:298-346): sizes without models pass;
    missing distance key -> KeyError; threshold ``>= round(n/2, 0)`` (banker's: 5 models -> 2)."""
    models = lr_models()
    if str(effective_size) not in models:
        return True
    votes = []
    model_list = models[str(effective_size)][str(pairwise_distance)]
    for c0, c1, b0, thr in model_list:
        value = 1 / (1 + math.e ** -(b0 + c0 * rmsd + c1 * orient))
        votes.append(value >= thr)
    return bool(sum(votes) >= round(len(model_list) / 2, 0))
