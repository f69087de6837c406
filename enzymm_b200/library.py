"""Template library compiler: ``Sequence[Template]`` -> the tables of ``emm_library_desc``.

This is the host half of what ``pyjess.Jess(templates)`` does when it is constructed
(reference call site ``enzymm/jess_run.py:800``) plus the per-residue constants EnzyMM derives
for its orientation filter (``enzymm/template.py:213-304``; ``jess_run.py:298-346, 499-520``).

Three things are decided here and nowhere else:

* **typing** (SURVEY.md 8c rules 2-3) -- ``type_match`` below is the data-driven predicate; it is
  evaluated once per (template typing key, query atom kind) and frozen into a bit matrix, so a
  different reading of ``match_mode`` is a one-line change;
* **typing classes** of query atoms -- two (residue name, atom name) kinds that no template atom
  can tell apart share a class; class 0 binds nothing and such atoms never reach the GPU search;
* **the search plan** -- which atom of each template residue leads and in which order residues
  are placed.  Any plan yields the same matches; a good one places rare residue types first.
"""
from __future__ import annotations

import hashlib
import json
import math
import os
import pickle
import threading
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .chem import BACKGROUND_PERCENT, RESIDUE_ATOMS
from .template_atoms import JessTemplate

_CLASS_LOCK = threading.RLock()      # guards the typing-class tables of every CompiledLibrary

__all__ = ["type_match", "CompiledLibrary", "load_lr_models", "MAX_TEMPLATE_ATOMS",
           "MAX_RESIDUES", "LR_MODELS"]

MAX_TEMPLATE_ATOMS = 32
MAX_RESIDUES = 10
LR_MODELS = 5
_DATA = Path(__file__).resolve().parent / "data"


MODE1_READINGS = ("N_or_O", "same_element", "exact_name", "N_O_or_S")


def type_match(match_mode: int, residue_names: Sequence[str], atom_names: Sequence[str],
               query_residue: str, query_atom: str, mode1: str = "N_or_O") -> bool:
    """May a template atom with this typing bind a query atom named ``query_atom`` in a residue
    named ``query_residue``?  (SURVEY.md 8c rules 2-3; names are whitespace-stripped.)

    ``match_mode < 100`` restricts the residue name to ``residue_names``; ``match_mode % 100``
    selects the atom-name test: 0 exact, 3 same first character, 8 same second character,
    1 query atom is N or O (this last reading is not pinned by any reference vector).
    """
    if match_mode < 0:
        raise ValueError(f"unsupported match_mode {match_mode}")
    if match_mode < 100 and query_residue not in residue_names:
        return False
    kind = match_mode % 100
    if kind == 0:
        return query_atom in atom_names
    if kind == 1:
        # the one reading no reference vector pins (4005 of the 7607 shipped templates carry such an
        # atom); ``mode1`` selects an alternative so the exposure can be measured and a pin applied
        if mode1 == "N_or_O":
            return query_atom[:1] in ("N", "O")
        if mode1 == "same_element":
            return any(query_atom[:1] == n[:1] for n in atom_names)
        if mode1 == "exact_name":
            return query_atom in atom_names
        if mode1 == "N_O_or_S":
            return query_atom[:1] in ("N", "O", "S")
        raise ValueError(f"unknown reading of match_mode 1: {mode1!r}")
    if kind == 3:
        return any(query_atom[:1] == n[:1] for n in atom_names)
    if kind == 8:
        # absent second characters compare equal (both padded), as in the oracle's NUL padding
        return any(query_atom[1:2] == n[1:2] for n in atom_names)
    raise ValueError(f"unsupported match_mode {match_mode}")


def load_lr_models(path: Optional[Path] = None) -> Dict[str, Dict[str, List[Tuple[float, float, float, float]]]]:
    """``{size: {distance: [(coef_rmsd, coef_orient, intercept, threshold) x 5]}}`` from
    ``logistic_regression_models.json`` (``enzymm/jess_run.py:499-520``)."""
    raw = json.loads((path or _DATA / "logistic_regression_models.json").read_text())
    out: Dict[str, Dict[str, List[Tuple[float, float, float, float]]]] = {}
    for size, by_dist in raw["match_size"].items():
        out[size] = {}
        for dist, entry in by_dist["pairwise_distance"].items():
            out[size][dist] = [(m["coef"][0], m["coef"][1], m["intercept"], m["threshold"])
                               for m in entry["model_list"]]
    return out


def _canonical_dist(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """sqrt((dx*dx + dy*dy) + dz*dz) with separately rounded float64 operations -- the
    expression the oracle and the CUDA guard-band path use."""
    d = a - b
    return np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])


def chain_code(chain_id: str) -> int:
    b = chain_id.encode("ascii", "replace")[:2].ljust(2, b"\0")
    return b[0] | (b[1] << 8)


class CompiledLibrary:
    """Host tables for one template list + one threshold assignment.

    Parameters
    ----------
    templates : sequence of ``JessTemplate`` (EnzyMM ``Template`` objects carry residues and an
        ``effective_size`` and get orientation + logistic tables; plain ones do not)
    rmsd_threshold, distance_cutoff, max_dynamic_distance : scalar or per-template sequences
        (``Matcher._get_jess_parameters``, ``jess_run.py:724-736``)
    """

    def __init__(self, templates: Sequence[JessTemplate], rmsd_threshold, distance_cutoff,
                 max_dynamic_distance, lr_models: Optional[dict] = None, plan_order: str = "leaders_first",
                 mode1: str = "N_or_O"):
        if plan_order not in ("leaders_first", "leaders_first_greedy", "residue_major"):
            raise ValueError("plan_order must be 'leaders_first', 'leaders_first_greedy' or 'residue_major'")
        if mode1 not in MODE1_READINGS:
            raise ValueError(f"mode1 must be one of {MODE1_READINGS}")
        self.plan_order = plan_order
        self.mode1 = mode1
        self.templates = list(templates)
        if not self.templates:
            raise ValueError("cannot compile an empty template list")
        T = len(self.templates)
        self.lr_models = load_lr_models() if lr_models is None else lr_models
        self._set_threshold_arrays(rmsd_threshold, distance_cutoff, max_dynamic_distance)

        # ---- typing keys -> compat rows ----------------------------------------------------
        self.ttype_of_key: Dict[tuple, int] = {}
        self.keys: List[tuple] = []
        atom_ttype: List[List[int]] = []
        for t in self.templates:
            atoms = list(t)
            if not 0 < len(atoms) <= MAX_TEMPLATE_ATOMS:
                raise ValueError(f"template {t.id!r} has {len(atoms)} atoms; supported: 1..{MAX_TEMPLATE_ATOMS}")
            row = []
            for a in atoms:
                key = a.typing_key()
                tt = self.ttype_of_key.get(key)
                if tt is None:
                    tt = self.ttype_of_key[key] = len(self.keys)
                    self.keys.append(key)
                row.append(tt)
            atom_ttype.append(row)
        self.n_ttype = len(self.keys)
        if self.n_ttype > 65535:
            raise ValueError("too many distinct template typing keys")

        # ---- typing classes of query atoms ---------------------------------------------------
        self._class_of_kind: Dict[Tuple[str, str], int] = {}
        self._class_of_column: Dict[bytes, int] = {}
        self._columns: List[np.ndarray] = [np.zeros(self.n_ttype, dtype=bool)]  # class 0: binds nothing
        self._class_of_column[self._columns[0].tobytes()] = 0
        self.compat_dirty = True
        expected = np.zeros(self.n_ttype, dtype=np.float64)  # expected candidates per 100 residues
        for res, names in RESIDUE_ATOMS.items():
            for name in names:
                c = self.class_of(res, name)
                if c:
                    expected += self._columns[c] * BACKGROUND_PERCENT[res]
        self.expected_candidates = expected

        # ---- search plans --------------------------------------------------------------------
        atom_off = [0]
        pair_off = [0]
        xyz, weight, chain = [], [], []
        plan_atom, plan_ttype, plan_src, plan_anchor = [], [], [], []
        pair_dist: List[np.ndarray] = []
        leader_of_ttype: Dict[int, int] = {}
        self.leader_ttype: List[int] = []
        for t, ttypes in zip(self.templates, atom_ttype):
            atoms = list(t)
            m = len(atoms)
            coords = np.array([(a.x, a.y, a.z) for a in atoms], dtype=np.float64)
            order, src_of = self._plan(atoms, ttypes, coords, float(self.distance_cutoff[len(atom_off) - 1]))
            pos_of_atom = {a: k for k, a in enumerate(order)}
            for k, a in enumerate(order):
                plan_atom.append(a)
                plan_ttype.append(ttypes[a])
                lead_atom = src_of[a]
                if lead_atom is None:
                    tt = ttypes[a]
                    lid = leader_of_ttype.get(tt)
                    if lid is None:
                        lid = leader_of_ttype[tt] = len(self.leader_ttype)
                        self.leader_ttype.append(tt)
                    plan_src.append(-1 - lid)
                else:
                    plan_src.append(pos_of_atom[lead_atom])
            pc = coords[np.asarray(order)]
            tri = [_canonical_dist(pc[k], pc[j]) for k in range(m) for j in range(k)]
            for k in range(m):
                src = plan_src[len(plan_src) - m + k]
                if k == 0:
                    plan_anchor.append(0)
                elif src >= 0:
                    plan_anchor.append(src)                      # same-residue position: its leader
                else:                                            # leader: the nearest placed atom gives the thinnest shell
                    row = tri[k * (k - 1) // 2: k * (k - 1) // 2 + k]
                    plan_anchor.append(int(np.argmin(row)))
            pair_dist.append(np.asarray(tri, dtype=np.float64))
            pair_off.append(pair_off[-1] + m * (m - 1) // 2)
            xyz.append(coords)
            weight.extend(a.distance_weight for a in atoms)
            chain.extend(chain_code(a.chain_id) for a in atoms)
            atom_off.append(atom_off[-1] + m)
        if len(self.leader_ttype) > 1023:
            raise ValueError("too many distinct leader types (limit 1023)")
        self.atom_off = np.asarray(atom_off, dtype=np.int32)
        self.pair_off = np.asarray(pair_off, dtype=np.int64)
        self.xyz = np.ascontiguousarray(np.concatenate(xyz), dtype=np.float64)
        self.weight = np.asarray(weight, dtype=np.float64)
        self.chain = np.asarray(chain, dtype=np.uint16)
        self.plan_atom = np.asarray(plan_atom, dtype=np.uint8)
        self.plan_ttype = np.asarray(plan_ttype, dtype=np.uint16)
        self.plan_src = np.asarray(plan_src, dtype=np.int16)
        self.plan_anchor = np.asarray(plan_anchor, dtype=np.uint8)
        self.pair_dist = np.ascontiguousarray(np.concatenate(pair_dist) if pair_dist else np.zeros(0), dtype=np.float64)
        if self.pair_dist.size == 0:
            self.pair_dist = np.zeros(1, dtype=np.float64)
        self.leader_ttype_arr = np.asarray(self.leader_ttype, dtype=np.uint16)

        # ---- orientation + logistic filter tables --------------------------------------------
        self.n_residues = np.zeros(T, dtype=np.int32)
        self.orient_idx = np.zeros((T, MAX_RESIDUES, 2), dtype=np.uint8)
        self.orient_vec = np.zeros((T, MAX_RESIDUES, 3), dtype=np.float64)
        for i, t in enumerate(self.templates):
            residues = getattr(t, "residues", None)
            if not residues:
                continue
            if len(residues) > MAX_RESIDUES:
                raise ValueError(f"template {t.id!r} has {len(residues)} residues; supported: {MAX_RESIDUES}")
            if 3 * len(residues) != len(t):
                raise ValueError("template residues do not tile its atoms in triplets")
            self.n_residues[i] = len(residues)
            for r, res in enumerate(residues):
                self.orient_idx[i, r] = res.orientation_vector_indices
                v = res.orientation_vector
                self.orient_vec[i, r] = (v.x, v.y, v.z)
        self._build_lr_tables()

    # ------------------------------------------------------------------------------------------
    def _set_threshold_arrays(self, rmsd_threshold, distance_cutoff, max_dynamic_distance):
        T = len(self.templates)

        def per_template(v):
            if np.ndim(v) == 0:
                return [v] * T
            if len(v) != T:
                raise ValueError("per-template threshold list has the wrong length")
            return list(v)

        self.distance_values = per_template(distance_cutoff)   # kept verbatim: str() keys the LR table
        self.rmsd_threshold = np.asarray(per_template(rmsd_threshold), dtype=np.float64)
        self.distance_cutoff = np.asarray(self.distance_values, dtype=np.float64)
        self.max_dynamic_distance = np.asarray(per_template(max_dynamic_distance), dtype=np.float64)

    def _build_lr_tables(self):
        rows: Dict[Tuple[str, str], int] = {}
        table: List[List[Tuple[float, float, float, float]]] = []
        self.lr_index = np.full(len(self.templates), -1, dtype=np.int32)
        for i, t in enumerate(self.templates):
            size = getattr(t, "effective_size", None)
            if size is None or str(size) not in self.lr_models or self.n_residues[i] == 0:
                continue
            by_dist = self.lr_models[str(size)]
            dist_key = str(self.distance_values[i])
            if dist_key not in by_dist:
                self.lr_index[i] = -2   # host raises KeyError when such a hit is filtered (jess_run.py:339-342)
                continue
            models = by_dist[dist_key]
            if len(models) != LR_MODELS:
                raise ValueError(f"expected {LR_MODELS} logistic models per cell, got {len(models)}")
            row = rows.get((str(size), dist_key))
            if row is None:
                row = rows[(str(size), dist_key)] = len(table)
                table.append(models)
            self.lr_index[i] = row
        self.lr_table = np.asarray(table, dtype=np.float64).reshape(-1, LR_MODELS, 4) if table else np.zeros((1, LR_MODELS, 4))
        self.n_lr = len(table)

    def set_thresholds(self, rmsd_threshold, distance_cutoff, max_dynamic_distance):
        """New threshold triple for the same templates (what ``Jess.query`` takes per call)."""
        self._set_threshold_arrays(rmsd_threshold, distance_cutoff, max_dynamic_distance)
        self._build_lr_tables()

    # ------------------------------------------------------------------------------------------
    # ---- expected-work model behind the leader order ---------------------------------------------
    _NOMINAL_RESIDUES = 400.0        # structure size the order is tuned for (AlphaFold-sized chains)

    def _leader_sequence(self, leaders, ttypes, coords, delta: float):
        """Order in which the residue leaders are placed: the permutation with the least expected
        work under a uniform-density model (exhaustive with branch and bound; the greedy
        rarest-and-nearest order is the starting bound).  Any order gives the same matches."""
        G = len(leaders)
        nominal = self._NOMINAL_RESIDUES
        scale = nominal / 100.0
        n = [max(float(self.expected_candidates[ttypes[a]]) * scale, 0.25) for a in leaders]
        radius = (3.0 * nominal * 135.0 / (4.0 * math.pi)) ** (1.0 / 3.0)
        pts = coords[leaders]
        dist = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(axis=2))

        def shell(d: float) -> float:     # P(|query distance - d| <= delta) for two random atoms of a globule
            lo, hi = max(d - delta, 0.0), min(d + delta, 2.0 * radius)
            if hi <= lo:
                return 1e-6
            cdf = lambda x: (x / radius) ** 3 - 9.0 / 16.0 * (x / radius) ** 4 + (x / radius) ** 6 / 32.0
            return min(1.0, max(cdf(hi) - cdf(lo), 1e-6))

        f = [[shell(float(dist[i, j])) if i != j else 1.0 for j in range(G)] for i in range(G)]
        VALIDATE, LEVEL = 10.0, 600.0     # cost of one validation step / one level visit, in filter tests
        # (measured on the bench workload: 367-370 ms for VALIDATE 3..10, LEVEL 100..2000, 250..800 nominal
        # residues, against 377 ms for the greedy order -- the model is not sensitive to its constants)

        def step(seq, survivors, g):
            """(added cost, survivors after placing g) given the placed sequence."""
            k = len(seq)
            tests = survivors * n[g]
            anchor = min(seq, key=lambda j: dist[g, j])
            pushes = tests * f[g][anchor]
            p_all = 1.0
            for j in seq:
                p_all *= f[g][j]
            after = tests * p_all
            return tests + VALIDATE * pushes * k + LEVEL * min(1.0, pushes), after

        # greedy start: rarest first, then cheapest expected shell
        first = min(range(G), key=lambda g: (n[g], g))
        greedy, surv, best_cost = [first], n[first], 0.0
        rest = [g for g in range(G) if g != first]
        while rest:
            g = min(rest, key=lambda g: (n[g] * f[g][min(greedy, key=lambda j: dist[g, j])], g))
            c, surv = step(greedy, surv, g)
            best_cost += c
            greedy.append(g)
            rest.remove(g)
        best = list(greedy)
        if G <= 2 or G > 8:
            return best
        budget = [200000]

        def search(seq, used, survivors, cost):
            nonlocal best, best_cost
            if cost >= best_cost or budget[0] <= 0:
                return
            if len(seq) == G:
                best, best_cost = list(seq), cost
                return
            budget[0] -= 1
            cands = []
            for g in range(G):
                if not used[g]:
                    c, after = step(seq, survivors, g)
                    cands.append((c, g, after))
            for c, g, after in sorted(cands):
                used[g] = True
                seq.append(g)
                search(seq, used, after, cost + c)
                seq.pop()
                used[g] = False

        for g0 in sorted(range(G), key=lambda g: n[g]):
            used = [False] * G
            used[g0] = True
            search([g0], used, n[g0], 0.0)
        return best

    def _plan(self, atoms, ttypes, coords, delta: float = 2.0):
        """Choose leaders and placement order.  Returns (order, src_of) where ``order`` lists
        template-order atom indices in placement order and ``src_of[a]`` is the atom whose query
        residue atom ``a`` must share (None for leaders)."""
        groups: Dict[tuple, List[int]] = {}
        for i, a in enumerate(atoms):
            groups.setdefault((a.chain_id, a.residue_number), []).append(i)
        glist = list(groups.values())
        cost = self.expected_candidates
        leaders = [min(g, key=lambda i: (cost[ttypes[i]], i)) for g in glist]
        if self.plan_order == "leaders_first":
            seq = self._leader_sequence(leaders, ttypes, coords, delta)
        else:
            remaining = list(range(len(glist)))
            first = min(remaining, key=lambda g: (cost[ttypes[leaders[g]]], g))
            seq = [first]
            remaining.remove(first)
            while remaining:
                placed = coords[[leaders[g] for g in seq]]

                def score(g):
                    d = _canonical_dist(placed, coords[leaders[g]]).min()
                    return (cost[ttypes[leaders[g]]] * max(float(d), 3.0) ** 2, g)

                nxt = min(remaining, key=score)
                seq.append(nxt)
                remaining.remove(nxt)
        order: List[int] = []
        src_of: Dict[int, Optional[int]] = {}
        if self.plan_order in ("leaders_first", "leaders_first_greedy"):
            # every residue's leader first (cheap one-distance filters prune whole residues), then
            # the remaining atoms of each residue, which only have to be looked up inside it
            for g in seq:
                order.append(leaders[g])
                src_of[leaders[g]] = None
            for g in seq:
                for i in glist[g]:
                    if i != leaders[g]:
                        order.append(i)
                        src_of[i] = leaders[g]
            return order, src_of
        for g in seq:
            lead = leaders[g]
            order.append(lead)
            src_of[lead] = None
            for i in glist[g]:
                if i != lead:
                    order.append(i)
                    src_of[i] = lead
        return order, src_of

    # ---- typing classes ------------------------------------------------------------------------
    def class_of(self, residue_name: str, atom_name: str) -> int:
        """Typing class of one query atom kind (creates the class on first sight)."""
        kind = (residue_name, atom_name)
        c = self._class_of_kind.get(kind)
        if c is None:
            with _CLASS_LOCK:
                return self._new_kind(kind)
        return c

    def _new_kind(self, kind) -> int:
        residue_name, atom_name = kind
        c = self._class_of_kind.get(kind)
        if c is None:
            column = np.fromiter((type_match(k[0], k[1], k[2], residue_name, atom_name, self.mode1) for k in self.keys),
                                 dtype=bool, count=self.n_ttype)
            sig = column.tobytes()
            c = self._class_of_column.get(sig)
            if c is None:
                c = len(self._columns)
                if c >= 1024:
                    raise ValueError("more than 1024 typing classes")
                self._columns.append(column)
                self._class_of_column[sig] = c
                self.compat_dirty = True
            self._class_of_kind[kind] = c
        return c

    def classify(self, residue_names: np.ndarray, atom_names: np.ndarray) -> np.ndarray:
        """Vectorised ``class_of`` over NumPy string columns -> uint16 classes."""
        n = len(residue_names)
        if n == 0:
            return np.zeros(0, dtype=np.uint16)
        packed = np.char.add(np.char.add(residue_names.astype("U4"), "|"), atom_names.astype("U4"))
        kinds, inverse = np.unique(packed, return_inverse=True)
        classes = np.empty(len(kinds), dtype=np.uint16)
        for i, k in enumerate(kinds):
            res, _, name = str(k).partition("|")
            classes[i] = self.class_of(res, name)
        return classes[inverse]

    # ---- vectorised classification of packed atom kinds ----------------------------------------
    _HASH_BITS = 16
    _HASH_MULT = np.uint64(0x9E3779B97F4A7C15)

    def classify_keys(self, keys: np.ndarray) -> np.ndarray:
        """Typing classes for atom kinds packed as ``resname_u32 << 32 | name_u32`` (the bytes of
        the blank-stripped names, little endian).  A 64 K-slot hash table (kind -> class) answers
        with two gathers; kinds it has not seen, or that lost their slot to another kind, are
        resolved through ``class_of``."""
        table = self.__dict__.get("_kind_table")
        if table is None:
            size = 1 << self._HASH_BITS
            table = self.__dict__["_kind_table"] = (np.full(size, np.uint64(0xFFFFFFFFFFFFFFFF)), np.zeros(size, np.uint16), {})
        slot_key, slot_class, overflow = table
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        slots = ((keys * self._HASH_MULT) >> np.uint64(64 - self._HASH_BITS)).astype(np.intp)
        out = slot_class[slots]
        miss = slot_key[slots] != keys
        if miss.any():
            where = np.nonzero(miss)[0]
            missing, inverse = np.unique(keys[where], return_inverse=True)
            resolved = np.empty(len(missing), dtype=np.uint16)
            with _CLASS_LOCK:
                for i, k in enumerate(missing.tolist()):
                    c = overflow.get(k)
                    if c is None:
                        text = lambda v: int(v).to_bytes(4, "little").rstrip(b"\0").decode("ascii", "replace")
                        c = self.class_of(text(k >> 32), text(k & 0xFFFFFFFF))
                        slot = ((k * int(self._HASH_MULT)) & 0xFFFFFFFFFFFFFFFF) >> (64 - self._HASH_BITS)
                        if slot_key[slot] == np.uint64(0xFFFFFFFFFFFFFFFF):
                            slot_key[slot], slot_class[slot] = np.uint64(k), c
                        else:
                            overflow[k] = c          # slot taken by another kind: answered from here
                    resolved[i] = c
            out[where] = resolved[inverse]
        return out

    @property
    def n_classes(self) -> int:
        return len(self._columns)

    @property
    def class_words(self) -> int:
        return (self.n_classes + 31) // 32

    def take_compat(self):
        """``(class_words, matrix)`` of the current classes, clearing ``compat_dirty`` -- atomically
        with respect to ``class_of`` running on another thread (the ingest thread of
        ``Matcher.scan_files`` may meet new atom kinds while a batch is being uploaded)."""
        with _CLASS_LOCK:
            matrix = self.compat_matrix()
            self.compat_dirty = False
            return matrix.shape[1], matrix

    def compat_matrix(self) -> np.ndarray:
        """``uint32 [n_ttype, class_words]`` bit matrix: bit c of row r = class c binds ttype r."""
        columns = list(self._columns)
        cw = (len(columns) + 31) // 32
        cols = np.stack(columns, axis=1)                            # [n_ttype, n_classes]
        padded = np.zeros((self.n_ttype, cw * 32), dtype=bool)
        padded[:, :cols.shape[1]] = cols
        bits = padded.reshape(self.n_ttype, cw, 32)
        weights = (1 << np.arange(32, dtype=np.uint64)).astype(np.uint64)
        return np.ascontiguousarray((bits * weights).sum(axis=2).astype(np.uint32))

    def __len__(self):
        return len(self.templates)

    # ---- compiled-library cache: compile the 7607-file library once ---------------
    _CACHE_VERSION = 6

    @staticmethod
    def _digest(templates, rmsd_threshold, distance_cutoff, max_dynamic_distance, plan_order) -> str:
        """Everything the compiled tables depend on: the template atoms, the per-residue orientation
        constants EnzyMM derives from them, the thresholds, the plan order, the logistic-model file
        and the planner's background tables."""
        h = hashlib.blake2b(digest_size=16)
        h.update(repr((CompiledLibrary._CACHE_VERSION, plan_order, np.asarray(rmsd_threshold, dtype=np.float64).tolist(),
                       [str(d) for d in np.atleast_1d(np.asarray(distance_cutoff, dtype=object))],
                       np.asarray(max_dynamic_distance, dtype=np.float64).tolist())).encode())
        h.update((_DATA / "logistic_regression_models.json").read_bytes())
        h.update(repr((sorted(BACKGROUND_PERCENT.items()), sorted((k, tuple(v)) for k, v in RESIDUE_ATOMS.items()))).encode())
        for t in templates:
            orient = [(tuple(r.orientation_vector_indices), (r.orientation_vector.x, r.orientation_vector.y,
                                                             r.orientation_vector.z))
                      for r in (getattr(t, "residues", None) or [])]
            h.update(repr((getattr(t, "effective_size", None), [a._key() for a in t], orient)).encode())
        return h.hexdigest()

    @classmethod
    def cached(cls, templates: Sequence[JessTemplate], rmsd_threshold, distance_cutoff, max_dynamic_distance,
               cache_dir, plan_order: str = "leaders_first") -> "CompiledLibrary":
        """Compile, or reload the tables compiled earlier for exactly these templates and thresholds
        from ``cache_dir`` (keyed by a digest of the template atoms and parameters)."""
        templates = list(templates)
        cache_dir = Path(cache_dir)
        key = cls._digest(templates, rmsd_threshold, distance_cutoff, max_dynamic_distance, plan_order)
        path = cache_dir / f"emm_library_{key}.pkl"
        if path.exists():
            with open(path, "rb") as handle:
                state = _TablesUnpickler(handle).load()     # arrays and plain containers only: no code runs
            self = cls.__new__(cls)
            self.__dict__.update(state)
            self.templates = templates
            self.compat_dirty = True
            return self
        self = cls(templates, rmsd_threshold, distance_cutoff, max_dynamic_distance, plan_order=plan_order)
        cache_dir.mkdir(parents=True, exist_ok=True, mode=0o700)
        state = {k: v for k, v in self.__dict__.items() if k not in ("templates", "_kind_table")}
        tmp = path.with_suffix(f".{os.getpid()}.tmp")
        with open(os.open(tmp, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o600), "wb") as handle:
            pickle.dump(state, handle, protocol=pickle.HIGHEST_PROTOCOL)
        tmp.replace(path)
        return self


class _TablesUnpickler(pickle.Unpickler):
    """Unpickler for the compiled-library cache: NumPy arrays / scalars / dtypes and plain containers
    only.  A cache file is data written by ``CompiledLibrary.cached``; anything else in it (a global
    that could run code on load) is refused."""

    _ALLOWED = {
        ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
        ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
        ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy._core.numeric", "_frombuffer"),
        ("numpy.core.numeric", "_frombuffer"), ("collections", "OrderedDict"), ("builtins", "set"),
        ("builtins", "frozenset"), ("builtins", "slice"), ("builtins", "bytearray"), ("builtins", "complex"),
    }

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED or (module == "numpy.dtypes" and name.endswith("DType")):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"compiled-library cache refers to {module}.{name}: refusing to load it")
