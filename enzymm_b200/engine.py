"""ctypes binding of ``libenzymm_b200.so`` (C ABI: ``include/enzymm_b200.h``).

The product path has no CPU fallback: if the CUDA library is missing or no GPU is visible,
``Engine`` raises.  The oracle under ``oracle/`` is test infrastructure and is never imported
from here.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path
from typing import Optional

import numpy as np

from .library import CompiledLibrary, MAX_TEMPLATE_ATOMS

__all__ = ["Engine", "DeviceLibrary", "Session", "PackedBatch", "HIT_DTYPE", "EngineError",
           "library_path", "load_cdll", "HIT_OVERFLOW", "HIT_BORDERLINE", "HIT_PASS",
           "HIT_NO_MODEL", "HIT_ORIENTED"]

HIT_OVERFLOW, HIT_BORDERLINE, HIT_PASS, HIT_NO_MODEL, HIT_ORIENTED = 0x01, 0x02, 0x04, 0x08, 0x10

EMM_OK = 0
_STATUS = {-1: "EMM_ERR_INVALID", -2: "EMM_ERR_CUDA", -3: "EMM_ERR_NO_DEVICE", -4: "EMM_ERR_CAPACITY",
           -5: "EMM_ERR_INPUT", -6: "EMM_ERR_NOMEM"}

HIT_DTYPE = np.dtype([
    ("structure", np.int32), ("template_index", np.int32), ("n_complete", np.uint32),
    ("n_atoms", np.uint16), ("flags", np.uint16), ("rmsd", np.float64), ("orientation", np.float64),
    ("rot", np.float64, (9,)), ("qbar", np.float64, (3,)), ("tbar", np.float64, (3,)),
    ("atoms", np.int32, (MAX_TEMPLATE_ATOMS,)),
], align=True)

STATS_FIELDS = ("pairs", "sweeps", "dist_evals", "exact_rechecks", "complete", "kept_atoms",
                "staged_bytes", "global_blobs")


class EngineError(RuntimeError):
    """A negative ``emm_status``.  For ``EMM_ERR_INPUT`` raised by a download, ``hits`` holds the hit
    records of the batch's valid structures (they were searched) and ``bad_structures`` maps the index
    of every skipped structure to its status (1 residue order, 2 too many atoms, 3 residue too large)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"{_STATUS.get(status, status)}: {message}")
        self.status = status
        self.hits = None
        self.bad_structures = {}


class _LibraryDesc(ctypes.Structure):
    _fields_ = [
        ("n_templates", ctypes.c_int32), ("n_atoms", ctypes.c_int32),
        ("atom_off", ctypes.c_void_p), ("xyz", ctypes.c_void_p), ("weight", ctypes.c_void_p),
        ("chain", ctypes.c_void_p), ("plan_atom", ctypes.c_void_p), ("plan_ttype", ctypes.c_void_p),
        ("plan_src", ctypes.c_void_p), ("plan_anchor", ctypes.c_void_p), ("pair_off", ctypes.c_void_p),
        ("pair_dist", ctypes.c_void_p),
        ("n_ttype", ctypes.c_int32), ("class_words", ctypes.c_int32), ("compat", ctypes.c_void_p),
        ("n_leader", ctypes.c_int32), ("leader_ttype", ctypes.c_void_p),
        ("rmsd_threshold", ctypes.c_void_p), ("distance_cutoff", ctypes.c_void_p),
        ("max_dynamic_distance", ctypes.c_void_p),
        ("n_residues", ctypes.c_void_p), ("orient_idx", ctypes.c_void_p), ("orient_vec", ctypes.c_void_p),
        ("lr_index", ctypes.c_void_p), ("n_lr", ctypes.c_int32), ("lr_table", ctypes.c_void_p),
    ]


class _Batch(ctypes.Structure):
    _fields_ = [
        ("n_structures", ctypes.c_int32), ("n_atoms", ctypes.c_int64),
        ("atom_off", ctypes.c_void_p), ("xyz", ctypes.c_void_p), ("klass", ctypes.c_void_p),
        ("residue", ctypes.c_void_p), ("bfactor", ctypes.c_void_p), ("chain", ctypes.c_void_p),
        ("atom_id", ctypes.c_void_p),
    ]


class _QueryParams(ctypes.Structure):
    _fields_ = [
        ("max_candidates", ctypes.c_int64), ("ignore_chain", ctypes.c_int32),
        ("conservation_cutoff", ctypes.c_float), ("template_begin", ctypes.c_int32),
        ("template_end", ctypes.c_int32), ("skip_mode", ctypes.c_int32),
        ("reset_structure_state", ctypes.c_int32), ("force_prepare", ctypes.c_int32),
        ("cell_threshold", ctypes.c_int32), ("donate_after", ctypes.c_int32),
    ]


class _Stats(ctypes.Structure):
    _fields_ = [(name, ctypes.c_uint64) for name in STATS_FIELDS]


def library_path() -> Path:
    override = os.environ.get("EMM_LIBRARY")          # development: try another build of the same ABI
    return Path(override) if override else Path(__file__).resolve().parent / "libenzymm_b200.so"


_cdll = None


def load_cdll() -> ctypes.CDLL:
    """Load the in-tree CUDA library; fail loudly when it was not built (``__graft_entry__.build``)."""
    global _cdll
    if _cdll is None:
        path = library_path()
        if not path.exists():
            raise ImportError(f"{path} is missing: build it with `make -C enzymm_b200/csrc` "
                              "(enzymm_b200 has no CPU fallback)")
        lib = ctypes.CDLL(str(path))
        lib.emm_last_error.restype = ctypes.c_char_p
        for name in ("emm_abi_version", "emm_hit_size", "emm_device_count", "emm_library_create", "emm_library_set_compat",
                     "emm_library_set_thresholds", "emm_library_set_filter", "emm_session_create", "emm_session_upload",
                     "emm_session_run", "emm_session_download", "emm_session_last_launches",
                     "emm_session_structure_status",
                     "emm_session_kernel_ms", "emm_session_clear_timings", "emm_session_debug_counters",
                     "emm_stream_create", "emm_stream_destroy",
                     "emm_query_batch"):
            getattr(lib, name).restype = ctypes.c_int
        lib.emm_library_destroy.restype = None
        lib.emm_session_destroy.restype = None
        if lib.emm_abi_version() != 2:
            raise ImportError("libenzymm_b200.so ABI version mismatch")
        assert ctypes.sizeof(_Stats) == 64
        if lib.emm_hit_size() != HIT_DTYPE.itemsize:
            raise ImportError("emm_hit layout mismatch between header and binding")
        _cdll = lib
    return _cdll


def _check(rc: int):
    if rc != EMM_OK:
        raise EngineError(rc, load_cdll().emm_last_error().decode(errors="replace"))


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class PackedBatch:
    """SoA columns of a batch of query structures, as ``emm_batch`` wants them (host memory)."""

    def __init__(self, atom_off: np.ndarray, xyz: np.ndarray, klass: np.ndarray, residue: np.ndarray,
                 bfactor: Optional[np.ndarray] = None, chain: Optional[np.ndarray] = None,
                 atom_id: Optional[np.ndarray] = None):
        self.atom_off = np.ascontiguousarray(atom_off, dtype=np.int64)
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        self.klass = np.ascontiguousarray(klass, dtype=np.uint16)
        self.residue = np.ascontiguousarray(residue, dtype=np.int32)
        self.bfactor = None if bfactor is None else np.ascontiguousarray(bfactor, dtype=np.float32)
        self.chain = None if chain is None else np.ascontiguousarray(chain, dtype=np.uint16)
        self.atom_id = None if atom_id is None else np.ascontiguousarray(atom_id, dtype=np.int32)
        self.table = None              # tsv.TableColumns when the batch came from a native packer
        n = int(self.atom_off[-1]) if len(self.atom_off) else 0
        for name in ("klass", "residue", "bfactor", "chain", "atom_id"):
            col = getattr(self, name)
            if col is not None and len(col) != n:
                raise ValueError(f"column {name} has {len(col)} entries, expected {n}")
        if len(self.xyz) != n:
            raise ValueError("xyz has the wrong length")

    @property
    def n_structures(self) -> int:
        return len(self.atom_off) - 1

    @property
    def n_atoms(self) -> int:
        return int(self.atom_off[-1])

    def nbytes(self) -> int:
        return sum(a.nbytes for a in (self.atom_off, self.xyz, self.klass, self.residue, self.bfactor,
                                      self.chain, self.atom_id) if a is not None)

    def slice(self, lo: int, hi: int) -> "PackedBatch":
        a0, a1 = int(self.atom_off[lo]), int(self.atom_off[hi])
        cut = lambda c: None if c is None else c[a0:a1]
        return PackedBatch(self.atom_off[lo:hi + 1] - a0, self.xyz[a0:a1], self.klass[a0:a1],
                           self.residue[a0:a1], cut(self.bfactor), cut(self.chain), cut(self.atom_id))

    # packed binary cache: parse and classify once, reload at memory speed
    def save(self, path) -> None:
        """Write the batch as an uncompressed ``.npz``.  Typing classes are only meaningful for
        the compiled library that produced them, so store the library digest next to the file."""
        cols = {k: getattr(self, k) for k in ("atom_off", "xyz", "klass", "residue", "bfactor", "chain", "atom_id")
                if getattr(self, k) is not None}
        np.savez(path, **cols)

    @classmethod
    def load(cls, path) -> "PackedBatch":
        with np.load(path) as z:
            get = lambda k: z[k] if k in z.files else None
            return cls(z["atom_off"], z["xyz"], z["klass"], z["residue"], get("bfactor"), get("chain"), get("atom_id"))

    def as_struct(self) -> _Batch:
        return _Batch(self.n_structures, self.n_atoms, _p(self.atom_off), _p(self.xyz), _p(self.klass),
                      _p(self.residue), _p(self.bfactor), _p(self.chain), _p(self.atom_id))


class DeviceLibrary:
    """A ``CompiledLibrary`` resident on one GPU (``emm_library``)."""

    def __init__(self, compiled: CompiledLibrary, device: int = 0):
        self.compiled = compiled
        self.device = device
        self._lib = load_cdll()
        c = compiled
        self._compat = c.compat_matrix()
        self._keep = [self._compat, c.leader_ttype_arr]
        desc = _LibraryDesc(
            len(c.templates), int(c.atom_off[-1]), _p(c.atom_off), _p(c.xyz), _p(c.weight), _p(c.chain),
            _p(c.plan_atom), _p(c.plan_ttype), _p(c.plan_src), _p(c.plan_anchor), _p(c.pair_off), _p(c.pair_dist),
            c.n_ttype, c.class_words, _p(self._compat), len(c.leader_ttype), _p(c.leader_ttype_arr),
            _p(c.rmsd_threshold), _p(c.distance_cutoff), _p(c.max_dynamic_distance),
            _p(c.n_residues), _p(c.orient_idx), _p(c.orient_vec), _p(c.lr_index), c.n_lr, _p(c.lr_table))
        handle = ctypes.c_void_p()
        _check(self._lib.emm_library_create(ctypes.c_int(device), ctypes.byref(desc), ctypes.byref(handle)))
        self.handle = handle
        c.compat_dirty = False

    def sync_compat(self):
        """Push the typing matrix again if classification created new classes."""
        c = self.compiled
        if c.compat_dirty:
            words, self._compat = c.take_compat()
            _check(self._lib.emm_library_set_compat(self.handle, ctypes.c_int32(words), _p(self._compat)))

    def push_thresholds(self):
        c = self.compiled
        _check(self._lib.emm_library_set_thresholds(self.handle, _p(c.rmsd_threshold), _p(c.distance_cutoff),
                                                    _p(c.max_dynamic_distance)))
        _check(self._lib.emm_library_set_filter(self.handle, _p(c.lr_index), ctypes.c_int32(c.n_lr),
                                                _p(c.lr_table)))

    def close(self):
        if getattr(self, "handle", None):
            self._lib.emm_library_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Session:
    """Device buffers for batches up to a fixed size (``emm_session``)."""

    def __init__(self, library: DeviceLibrary, max_atoms: int, max_structures: int, hit_capacity: int = 0):
        self.library = library
        self._lib = load_cdll()
        self.max_atoms, self.max_structures = max(1, int(max_atoms)), max(1, int(max_structures))
        self.hit_capacity = int(hit_capacity or max(1024, 64 * max_structures))
        handle = ctypes.c_void_p()
        _check(self._lib.emm_session_create(library.handle, ctypes.c_int64(max(1, int(max_atoms))),
                                            ctypes.c_int32(max(1, int(max_structures))),
                                            ctypes.c_int64(self.hit_capacity), ctypes.byref(handle)))
        self.handle = handle
        self._hits = np.zeros(self.hit_capacity, dtype=HIT_DTYPE)
        self._batch = None

    def upload(self, batch: PackedBatch, stream: int = 0):
        self.library.sync_compat()
        self._batch = batch            # keep host buffers alive until the copy is consumed
        st = batch.as_struct()
        _check(self._lib.emm_session_upload(self.handle, ctypes.byref(st), ctypes.c_void_p(stream)))

    def run(self, *, max_candidates: int = 10000, ignore_chain: bool = True, conservation_cutoff: float = 0.0,
            template_begin: int = 0, template_end: int = 0, skip_mode: int = 0, reset: bool = True,
            force_prepare: bool = False, cell_threshold: int = 0, donate_after: int = 0, stream: int = 0):
        q = _QueryParams(int(max_candidates or 0), 1 if ignore_chain else 0, float(conservation_cutoff or 0.0),
                         int(template_begin), int(template_end), int(skip_mode), 1 if reset else 0,
                         1 if force_prepare else 0, int(cell_threshold), int(donate_after))
        _check(self._lib.emm_session_run(self.handle, ctypes.byref(q), ctypes.c_void_p(stream)))

    def download(self, stream: int = 0, with_stats: bool = False, out: Optional[np.ndarray] = None):
        """Hits of the last run(s), sorted by (structure, template).  ``out``: a HIT_DTYPE array to
        receive them in place (e.g. a slice of one large pinned buffer) instead of a fresh copy."""
        n = ctypes.c_int64(0)
        stats = _Stats()
        if out is not None:
            if out.dtype != HIT_DTYPE or not out.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be a contiguous HIT_DTYPE array")
            rc = self._lib.emm_session_download(self.handle, _p(out), ctypes.c_int64(len(out)), ctypes.byref(n),
                                                ctypes.byref(stats), ctypes.c_void_p(stream))
            _check(rc)
            return (out[:n.value], {k: int(getattr(stats, k)) for k in STATS_FIELDS}) if with_stats else out[:n.value]
        rc = self._lib.emm_session_download(self.handle, _p(self._hits), ctypes.c_int64(self.hit_capacity),
                                            ctypes.byref(n), ctypes.byref(stats), ctypes.c_void_p(stream))
        if rc == -5:
            # structures that break the input contract were skipped; everything else was searched
            message = self._lib.emm_last_error().decode(errors="replace")
            exc = EngineError(rc, message)
            exc.hits = self._hits[:n.value].copy()
            status = self.structure_status()
            exc.bad_structures = {int(i): int(status[i]) for i in np.nonzero(status)[0]}
            raise exc
        _check(rc)
        hits = self._hits[:n.value].copy()
        if with_stats:
            return hits, {k: int(getattr(stats, k)) for k in STATS_FIELDS}
        return hits

    def structure_status(self) -> np.ndarray:
        """Per-structure outcome of the prepare pass of the uploaded batch (0 = searched)."""
        n = self._batch.n_structures if self._batch is not None else 0
        out = np.zeros(max(n, 1), dtype=np.int32)
        _check(self._lib.emm_session_structure_status(self.handle, _p(out), ctypes.c_int32(len(out))))
        return out[:n]

    @property
    def last_launches(self) -> int:
        return int(self._lib.emm_session_last_launches(self.handle))

    def kernel_ms(self, which: str):
        """Device durations (ms) of the 'prepare' / 'search' launches since ``clear_timings`` (CUDA
        events recorded on the launching stream; call after synchronising it)."""
        idx = {"prepare": 0, "search": 1}[which]
        count = ctypes.c_int(0)
        buf = np.zeros(4096, dtype=np.float32)
        _check(self._lib.emm_session_kernel_ms(self.handle, ctypes.c_int(idx), _p(buf), ctypes.c_int(len(buf)),
                                               ctypes.byref(count)))
        return buf[:count.value].astype(float).tolist()

    def debug_counters(self):
        buf = np.zeros(128, dtype=np.uint64)
        _check(self._lib.emm_session_debug_counters(self.handle, _p(buf)))
        return buf

    def clear_timings(self):
        _check(self._lib.emm_session_clear_timings(self.handle))

    def close(self):
        if getattr(self, "handle", None):
            self._lib.emm_session_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """Convenience front end: one compiled library on one device + a growable session."""

    def __init__(self, compiled: CompiledLibrary, device: int = 0):
        lib = load_cdll()
        if lib.emm_device_count() <= 0:
            raise EngineError(-3, "no CUDA device visible: enzymm_b200 has no CPU fallback")
        self.device_library = DeviceLibrary(compiled, device)
        self.compiled = compiled
        self.device = device
        self._session: Optional[Session] = None
        self._cap = (0, 0, 0)

    def new_stream(self) -> int:
        """A non-blocking CUDA stream on this engine's device (handle for the ``stream=`` arguments)."""
        st = ctypes.c_void_p()
        _check(load_cdll().emm_stream_create(ctypes.c_int(self.device), ctypes.byref(st)))
        return int(st.value or 0)

    def free_stream(self, stream: int) -> None:
        if stream:
            _check(load_cdll().emm_stream_destroy(ctypes.c_int(self.device), ctypes.c_void_p(stream)))

    def session_for(self, n_atoms: int, n_structures: int, hit_capacity: int = 0) -> Session:
        need_hits = hit_capacity or max(1024, 64 * n_structures)
        if (self._session is None or n_atoms > self._cap[0] or n_structures > self._cap[1]
                or need_hits > self._cap[2]):
            if self._session is not None:
                self._session.close()
            cap = (max(n_atoms, self._cap[0]), max(n_structures, self._cap[1]), max(need_hits, self._cap[2]))
            self._session = Session(self.device_library, cap[0], cap[1], cap[2])
            self._cap = cap
        return self._session

    def query(self, batch: PackedBatch, **params):
        """upload + run + download; on EMM_ERR_CAPACITY the hit buffer is enlarged and the batch re-run."""
        with_stats = params.pop("with_stats", False)
        hit_capacity = 0
        for _ in range(4):
            sess = self.session_for(batch.n_atoms, batch.n_structures, hit_capacity)
            sess.upload(batch)
            sess.run(**params)
            try:
                return sess.download(with_stats=with_stats)
            except EngineError as exc:
                if exc.status != -4:
                    raise
                hit_capacity = 4 * sess.hit_capacity
        raise EngineError(-4, "hit buffer kept overflowing")

    def close(self):
        if self._session is not None:
            self._session.close()
            self._session = None
        self.device_library.close()
