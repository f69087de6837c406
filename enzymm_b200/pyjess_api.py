"""The ``pyjess`` names EnzyMM imports, backed by the CUDA engine.

Drop-in surface: ``Jess(templates).query(molecule, rmsd_threshold,
distance_cutoff, max_dynamic_distance, max_candidates=..., best_match=..., ignore_chain=...)``
returning an iterator of ``Hit`` (reference call site ``enzymm/jess_run.py:800-811`` and
``tests/test_jess_run.py:32-39``), with ``Hit.rmsd / .log_evalue / .template / .atoms(transform)
/ .molecule(transform)`` as consumed at ``jess_run.py:148-182, 243-262, 317, 357, 397, 485``.

There is no CPU path: constructing the engine without the CUDA library or a GPU raises.
"""
from __future__ import annotations

import math
import threading
from collections import OrderedDict
from typing import Iterator, List, Optional, Sequence

import numpy as np

from . import __version__  # noqa: F401  (pyjess.__version__ is printed at jess_run.py:659)
from .engine import Engine, HIT_BORDERLINE, HIT_NO_MODEL, HIT_OVERFLOW, HIT_PASS
from .library import CompiledLibrary
from .packing import pack_molecules
from .structures import Atom, Molecule
from .template_atoms import JessTemplate as Template
from .template_atoms import TemplateAtom

__all__ = ["Atom", "Molecule", "TemplateAtom", "Template", "Jess", "Query", "Hit"]


class Hit:
    """One template matched onto one molecule (mirror of ``pyjess.Hit``).

    ``log_evalue`` is NaN: its formula lives in the un-vendored Jess C source, only two values
    are pinned by the reference tests and it feeds no decision (SURVEY.md 8c "parity unpinned").
    """

    __slots__ = ("rmsd", "log_evalue", "template", "_molecule", "_record", "_decoded",
                 "orientation", "flags", "n_complete", "template_index", "structure_index")

    def __init__(self, record: np.void, template, molecule: Molecule, scalars=None):
        """``record``: one ``emm_hit`` row.  ``scalars`` (optional) = that row's (rmsd, orientation,
        flags, n_complete, template_index, structure) as Python numbers, for callers that convert
        whole columns at once (``Matcher._assemble``); the matched atoms and the transform are
        decoded from the record when first asked for."""
        if scalars is None:
            scalars = (float(record["rmsd"]), float(record["orientation"]), int(record["flags"]),
                       int(record["n_complete"]), int(record["template_index"]), int(record["structure"]))
        self.rmsd, self.orientation, self.flags, self.n_complete, self.template_index, self.structure_index = scalars
        self.log_evalue = math.nan
        self.template = template
        self._molecule = molecule
        self._record = record
        self._decoded = None

    def _decode(self):
        d = self._decoded
        if d is None:
            r = self._record
            n = int(r["n_atoms"])
            d = self._decoded = (np.asarray(r["atoms"][:n], dtype=np.int64),
                                 np.asarray(r["rot"], dtype=np.float64).reshape(3, 3).copy(),
                                 np.asarray(r["qbar"], dtype=np.float64).copy(),
                                 np.asarray(r["tbar"], dtype=np.float64).copy())
        return d

    @property
    def _atom_idx(self):
        return self._decode()[0]

    @property
    def _rot(self):
        return self._decode()[1]

    @property
    def _qbar(self):
        return self._decode()[2]

    @property
    def _tbar(self):
        return self._decode()[3]

    # device-side verdicts ----------------------------------------------------------------------
    @property
    def device_pass(self) -> bool:
        """``Match.predicted_correct`` as decided by the fused filter on the GPU."""
        return bool(self.flags & HIT_PASS)

    @property
    def overflow(self) -> bool:
        return bool(self.flags & HIT_OVERFLOW)

    @property
    def borderline(self) -> bool:
        return bool(self.flags & HIT_BORDERLINE)

    @property
    def missing_model(self) -> bool:
        return bool(self.flags & HIT_NO_MODEL)

    @property
    def atom_indices(self) -> List[int]:
        """Matched query atom indices in template atom order."""
        return [int(i) for i in self._atom_idx]

    # pyjess surface -----------------------------------------------------------------------------
    def _transform(self, xyz: np.ndarray) -> np.ndarray:
        # q' = R (q - qbar) + tbar : query frame -> template frame (SURVEY 8c rule 9)
        return (xyz - self._qbar) @ self._rot.T + self._tbar

    def atoms(self, transform: bool = True) -> List[Atom]:
        """Matched query atoms in template order; ``transform=True`` -> template frame."""
        mol = self._molecule
        out = []
        xyz = mol.xyz[self._atom_idx]
        if transform:
            xyz = self._transform(xyz)
        for row, i in zip(xyz, self._atom_idx):
            atom = mol.atom(int(i))
            atom.x, atom.y, atom.z = float(row[0]), float(row[1]), float(row[2])
            out.append(atom)
        return out

    def molecule(self, transform: bool = False) -> Molecule:
        if not transform:
            return self._molecule
        return self._molecule.with_xyz(self._transform(self._molecule.xyz))

    def __repr__(self):
        return f"Hit(template={getattr(self.template, 'id', None)!r}, rmsd={self.rmsd:.4f})"


class Query:
    """Iterator over the hits of one ``Jess.query`` call (mirror of ``pyjess.Query``)."""

    def __init__(self, hits: List[Hit], molecule: Molecule, rmsd_threshold: float, distance_cutoff: float,
                 max_dynamic_distance: float, max_candidates, best_match: bool, ignore_chain: bool):
        self._hits = hits
        self._pos = 0
        self.molecule = molecule
        self.rmsd_threshold = rmsd_threshold
        self.distance_cutoff = distance_cutoff
        self.max_dynamic_distance = max_dynamic_distance
        self.max_candidates = max_candidates
        self.best_match = best_match
        self.ignore_chain = ignore_chain

    def __iter__(self) -> Iterator[Hit]:
        return self

    def __next__(self) -> Hit:
        if self._pos >= len(self._hits):
            raise StopIteration
        hit = self._hits[self._pos]
        self._pos += 1
        return hit


class _EngineEntry:
    """One cached engine: its own compiled library, device tables, session and lock."""

    __slots__ = ("engine", "templates", "lock", "users", "evicted")

    def __init__(self, engine: Engine, templates: Sequence[Template]):
        self.engine = engine
        self.templates = list(templates)     # keeps the templates alive so their ids stay unique
        self.lock = threading.Lock()         # one query at a time per engine (its session is not re-entrant)
        self.users = 0
        self.evicted = False


_ENGINE_CACHE: "OrderedDict[tuple, _EngineEntry]" = OrderedDict()
_ENGINE_CACHE_SIZE = 16
_CACHE_LOCK = threading.Lock()


def _acquire_engine(templates: Sequence[Template], device: int, rmsd_threshold: float, distance_cutoff: float,
                    max_dynamic_distance: float) -> _EngineEntry:
    """EnzyMM rebuilds ``Jess(templates)`` on every call (jess_run.py:800) and calls ``query`` from a
    ``ThreadPool`` (jess_run.py:919-921, 976-978); compiling and uploading a library per call would
    dominate, so engines are cached -- keyed by template identity AND the threshold triple, so a
    cached engine's device tables are never rewritten by a query and concurrent calls with other
    thresholds (other size groups) get their own engine."""
    key = (device, tuple(id(t) for t in templates), float(rmsd_threshold), str(distance_cutoff),
           float(max_dynamic_distance))
    with _CACHE_LOCK:
        entry = _ENGINE_CACHE.get(key)
        if entry is not None:
            _ENGINE_CACHE.move_to_end(key)
        else:
            compiled = CompiledLibrary(templates, rmsd_threshold, distance_cutoff, max_dynamic_distance)
            entry = _ENGINE_CACHE[key] = _EngineEntry(Engine(compiled, device), templates)
            while len(_ENGINE_CACHE) > _ENGINE_CACHE_SIZE:
                _, old = _ENGINE_CACHE.popitem(last=False)
                old.evicted = True
                if old.users == 0:
                    old.engine.close()
        entry.users += 1
        return entry


def _release_engine(entry: _EngineEntry) -> None:
    with _CACHE_LOCK:
        entry.users -= 1
        if entry.evicted and entry.users == 0:    # evicted while in use: the last user closes it
            entry.engine.close()


def clear_engine_cache() -> None:
    """Release every cached engine (device libraries and sessions) that is not in use."""
    with _CACHE_LOCK:
        while _ENGINE_CACHE:
            _, old = _ENGINE_CACHE.popitem(last=False)
            old.evicted = True
            if old.users == 0:
                old.engine.close()


def _device_query(templates: Sequence[Template], device: int, molecule: Molecule, rmsd_threshold: float,
                  distance_cutoff: float, max_dynamic_distance: float, max_candidates: int,
                  ignore_chain: bool) -> np.ndarray:
    """Hit records of one molecule against ``templates`` (re-entrant: safe to call from many threads)."""
    entry = _acquire_engine(templates, device, rmsd_threshold, distance_cutoff, max_dynamic_distance)
    try:
        with entry.lock:
            batch = pack_molecules([molecule], entry.engine.compiled)
            return entry.engine.query(batch, max_candidates=max_candidates, ignore_chain=ignore_chain)
    finally:
        _release_engine(entry)


class Jess:
    """A set of templates to query molecules against (mirror of ``pyjess.Jess``)."""

    def __init__(self, templates: Sequence[Template] = (), device: int = 0):
        self._templates = list(templates)
        self._device = device

    def __len__(self):
        return len(self._templates)

    def __iter__(self):
        return iter(self._templates)

    def __getitem__(self, i):
        return self._templates[i]

    def query(self, molecule: Molecule, rmsd_threshold: float, distance_cutoff: float,
              max_dynamic_distance: float, *, max_candidates: Optional[int] = None,
              ignore_chain: bool = False, best_match: bool = False) -> Query:
        """Match every template against ``molecule``; yields at most one ``Hit`` per template.

        Only ``best_match=True`` -- the one mode EnzyMM uses (jess_run.py:809) -- is implemented.
        """
        if not best_match:
            raise NotImplementedError(
                "enzymm_b200 implements best_match=True only (the mode EnzyMM uses, jess_run.py:809)")
        hits: List[Hit] = []
        if self._templates and len(molecule):
            # pyjess's own default when max_candidates is None is 1000 (comment at jess_run.py:796)
            cap = 1000 if max_candidates is None else int(max_candidates)
            records = _device_query(self._templates, self._device, molecule, rmsd_threshold, distance_cutoff,
                                    max_dynamic_distance, cap, ignore_chain)
            hits = [Hit(r, self._templates[int(r["template_index"])], molecule) for r in records]
        return Query(hits, molecule, rmsd_threshold, distance_cutoff, max_dynamic_distance,
                     max_candidates, best_match, ignore_chain)
