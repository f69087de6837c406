"""Query-structure containers: the ``pyjess.Atom`` / ``pyjess.Molecule`` surface EnzyMM uses.

This is the host-side input stage of the hot path.  The reference
gets these two types from the un-vendored ``pyjess`` wheel (call sites
``enzymm/jess_run.py:538-548`` for ``Molecule.load`` / ``.conserved``, ``jess_run.py:125-145``
for the ``Atom`` attribute surface, ``jess_run.py:485-496`` for ``len()`` / iteration).  Here a
molecule is a set of NumPy columns (SoA) so that a batch of molecules can be packed for upload
without touching per-atom Python objects; ``Atom`` views are materialised lazily.

Behaviour pinned by the reference's goldens (SURVEY.md §8c rule 1):
  * ``ATOM`` *and* ``HETATM`` records, in file order, up to the first ``ENDMDL``
    (3184 + 155 = 3339 atoms for ``1AMY.pdb``, ``tests/test_jess_run.py:75``);
  * ``conserved(c)`` returns a NEW molecule keeping atoms with ``temperature_factor >= c``
    (494 residues of the AlphaFold fixture at cutoff 80, ``tests/test_jess_run.py:276,326-329``);
  * the molecule id defaults to the HEADER idCode (``"1AMY"``, ``tests/test_jess_run.py:62``).
"""
from __future__ import annotations

import ctypes
import hashlib
import io
import os
from pathlib import Path
from typing import IO, Iterable, Iterator, List, Optional, Sequence, Union

import numpy as np

__all__ = ["Atom", "Molecule", "load_many"]

_STR4 = "U4"


class Atom:
    """One query atom (mirror of ``pyjess.Atom``; attribute names as used at
    ``enzymm/jess_run.py:125-145``)."""

    __slots__ = (
        "serial", "name", "altloc", "residue_name", "chain_id", "residue_number",
        "insertion_code", "x", "y", "z", "occupancy", "temperature_factor",
        "segment", "element", "charge",
    )

    def __init__(
        self,
        *,
        serial: int,
        name: str,
        residue_name: str,
        chain_id: str,
        residue_number: int,
        x: float,
        y: float,
        z: float,
        altloc: str = " ",
        insertion_code: str = " ",
        occupancy: float = 0.0,
        temperature_factor: float = 0.0,
        segment: str = "",
        element: str = "",
        charge: int = 0,
    ):
        self.serial = int(serial)
        self.name = str(name)
        self.altloc = altloc
        self.residue_name = str(residue_name)
        self.chain_id = str(chain_id)
        self.residue_number = int(residue_number)
        self.insertion_code = insertion_code
        self.x = float(x)
        self.y = float(y)
        self.z = float(z)
        self.occupancy = float(occupancy)
        self.temperature_factor = float(temperature_factor)
        self.segment = str(segment)
        self.element = str(element)
        self.charge = int(charge)

    def _key(self):
        return tuple(getattr(self, s) for s in self.__slots__)

    def __eq__(self, other):
        if not isinstance(other, Atom):
            return NotImplemented
        return self._key() == other._key()

    def __ne__(self, other):
        if not isinstance(other, Atom):
            return NotImplemented
        return self._key() != other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return (f"Atom(serial={self.serial}, name={self.name!r}, residue_name={self.residue_name!r}, "
                f"chain_id={self.chain_id!r}, residue_number={self.residue_number}, "
                f"x={self.x}, y={self.y}, z={self.z})")


_COLUMNS = (
    ("serial", np.int32), ("name", _STR4), ("altloc", "U1"), ("residue_name", _STR4),
    ("chain_id", "U2"), ("residue_number", np.int32), ("insertion_code", "U1"),
    ("occupancy", np.float64), ("temperature_factor", np.float64), ("segment", _STR4),
    ("element", "U2"), ("charge", np.int8),
)


_STR_WIDTH = {"name": 4, "altloc": 1, "residue_name": 4, "chain_id": 2, "insertion_code": 1, "segment": 4, "element": 2}


class _Columns(dict):
    """Column store of a ``Molecule``.  String columns that arrive from the native reader stay
    fixed-width byte matrices (``raw[name]``: uint8 ``[n, width]``, blank-stripped, NUL padded) and
    become NumPy unicode arrays only when somebody asks for them: a screening run touches the
    strings of a handful of matched atoms, not of every atom it read."""

    __slots__ = ("raw",)

    def __init__(self, converted=(), raw=None):
        super().__init__(converted)
        self.raw = dict(raw or {})

    def __missing__(self, key):
        r = self.raw.get(key)
        if r is None:
            raise KeyError(key)
        w = r.shape[1]
        arr = np.ascontiguousarray(r).reshape(-1).view(f"S{w}").astype(f"U{w}") if len(r) else np.zeros(0, dtype=f"U{w}")
        self[key] = arr
        return arr

    def names(self):
        return [k for k, _ in _COLUMNS if k in self or k in self.raw]

    def take(self, keep, copy: bool = False) -> "_Columns":
        pick = (lambda v: v[keep].copy()) if copy else (lambda v: v[keep])
        return _Columns({k: pick(v) for k, v in dict.items(self)}, {k: pick(v) for k, v in self.raw.items()})

    def text_at(self, key: str, i: int) -> str:
        """One string cell without converting the column."""
        if key in self:
            return str(dict.__getitem__(self, key)[i])
        return bytes(self.raw[key][i]).rstrip(b"\0").decode("ascii", "replace")

    def canonical_bytes(self, key: str) -> bytes:
        """Representation-independent bytes of a column (for hashing / equality)."""
        w = _STR_WIDTH.get(key)
        if w is None:
            return dict.__getitem__(self, key).tobytes()
        if key in self.raw:
            return np.ascontiguousarray(self.raw[key]).tobytes()
        return self.packed_u32(key).astype({4: "<u4", 2: "<u2", 1: "u1"}[w]).tobytes()

    def packed_u32(self, key: str) -> np.ndarray:
        """String column as one little-endian uint32 per cell (bytes of the blank-stripped text)."""
        w = _STR_WIDTH[key]
        if key in self.raw:
            r = self.raw[key]
            out = np.zeros(len(r), dtype=np.uint32)
            for j in range(w):
                out |= r[:, j].astype(np.uint32) << np.uint32(8 * j)
            return out
        arr = dict.__getitem__(self, key)
        if not len(arr):
            return np.zeros(0, dtype=np.uint32)
        width = arr.dtype.itemsize // 4
        code = np.ascontiguousarray(arr).view(np.uint32).reshape(len(arr), width)
        if (code > 255).any():
            code = np.where(code > 255, ord("?"), code)
        out = np.zeros(len(arr), dtype=np.uint32)
        for j in range(min(w, width)):
            out |= code[:, j].astype(np.uint32) << np.uint32(8 * j)
        return out


def _parse_pdb_text(lines: Iterable[str]):
    """Fixed-column PDB reader (ATOM/HETATM up to first ENDMDL)."""
    cols = {k: [] for k, _ in _COLUMNS}
    xs: List[float] = []
    ys: List[float] = []
    zs: List[float] = []
    header_id: Optional[str] = None
    for line in lines:
        rec = line[:6]
        if rec == "ATOM  " or rec == "HETATM":
            line = line.rstrip("\r\n")
            if len(line) < 54:
                raise ValueError(f"truncated PDB coordinate record: {line!r}")
            if len(line) < 80:
                line = line.ljust(80)
            try:
                cols["serial"].append(int(line[6:11]))
                cols["residue_number"].append(int(line[22:26]))
                xs.append(float(line[30:38]))
                ys.append(float(line[38:46]))
                zs.append(float(line[46:54]))
            except ValueError as exc:
                raise ValueError(f"malformed PDB coordinate record: {line!r}") from exc
            occ = line[54:60].strip()
            bfac = line[60:66].strip()
            cols["occupancy"].append(float(occ) if occ else 0.0)
            cols["temperature_factor"].append(float(bfac) if bfac else 0.0)
            cols["name"].append(line[12:16].strip())
            cols["altloc"].append(line[16])
            cols["residue_name"].append(line[17:20].strip())
            cols["chain_id"].append(line[20:22].strip())
            cols["insertion_code"].append(line[26])
            cols["segment"].append(line[72:76].strip())
            cols["element"].append(line[76:78].strip())
            chg = line[78:80].strip()
            charge = 0
            if chg and chg[0].isdigit():
                charge = int(chg[0]) * (-1 if chg.endswith("-") else 1)
            cols["charge"].append(charge)
        elif rec == "ENDMDL":
            break
        elif rec == "HEADER" and header_id is None:
            code = line[62:66].strip()
            header_id = code or None
    arrays = {k: np.asarray(cols[k], dtype=dt) for k, dt in _COLUMNS}
    xyz = np.empty((len(xs), 3), dtype=np.float64)
    xyz[:, 0] = xs
    xyz[:, 1] = ys
    xyz[:, 2] = zs
    return arrays, xyz, header_id


# ---- native ingest (enzymm_b200/csrc/emm_pdb.cpp through the C ABI) -----------------------------
_native = None


def _native_lib():
    """libenzymm_b200.so for its host-only PDB entry points (they need no GPU)."""
    global _native
    if _native is None:
        override = os.environ.get("EMM_LIBRARY")          # development: another build of the same ABI
        path = Path(override) if override else Path(__file__).resolve().parent / "libenzymm_b200.so"
        if not path.exists():
            raise ImportError(f"{path} is missing: build it with `make -C enzymm_b200/csrc`")
        lib = ctypes.CDLL(str(path))
        lib.emm_pdb_last_error.restype = ctypes.c_char_p
        lib.emm_pdb_batch_free.argtypes = [ctypes.c_void_p]
        lib.emm_pdb_batch_free.restype = None
        lib.emm_pdb_batch_file_message.argtypes = [ctypes.c_void_p, ctypes.c_int32]
        lib.emm_pdb_batch_file_message.restype = ctypes.c_char_p
        _native = lib
    return _native


class _NativeBatch:
    """Owns one ``emm_pdb_batch`` handle; freed when the last array viewing it goes away."""

    def __init__(self, lib, handle):
        self._lib, self._handle = lib, handle

    def __del__(self):
        if self._handle:
            self._lib.emm_pdb_batch_free(self._handle)
            self._handle = None


def _native_view(owner: "_NativeBatch", ptr, dtype, count: int) -> np.ndarray:
    """NumPy view of ``count`` items at ``ptr`` inside a native batch (array -> buffer -> owner keeps it alive)."""
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
    buf._owner = owner
    return np.frombuffer(buf, dtype=dtype, count=count)


class _PdbColumns(ctypes.Structure):
    _fields_ = [("n_files", ctypes.c_int32), ("n_atoms", ctypes.c_int64), ("atom_off", ctypes.c_void_p)] + \
               [(k, ctypes.c_void_p) for k in ("serial", "name", "altloc", "resname", "chain", "resnum", "icode",
                                               "xyz", "occupancy", "bfactor", "segment", "element", "charge",
                                               "header_id")]


def _fixed_to_str(raw: np.ndarray, width: int) -> np.ndarray:
    """[n*width] NUL padded bytes -> NumPy unicode column."""
    return raw.view(f"S{width}").astype(f"U{width}") if raw.size else np.zeros(0, dtype=f"U{max(width, 1)}")


def _columns_from_native(n, serial, name, altloc, resname, chain, resnum, icode, occ, bfac, segment, element, charge):
    raw = {"name": name.reshape(-1, 4), "altloc": altloc.reshape(-1, 1), "residue_name": resname.reshape(-1, 4),
           "chain_id": chain.reshape(-1, 2), "insertion_code": icode.reshape(-1, 1), "segment": segment.reshape(-1, 4),
           "element": element.reshape(-1, 2)}
    return _Columns({"serial": serial, "residue_number": resnum, "occupancy": occ, "temperature_factor": bfac,
                     "charge": charge}, raw)


CIF_AUTHOR = 1        # EMM_PDB_CIF_AUTHOR: read the auth_* identifiers of an mmCIF file before the label_* ones


def _parse_pdb_native(data: bytes, flags: int = 0):
    """One PDB or mmCIF text through emm_pdb_parse_ex; same columns as ``_parse_pdb_text``."""
    lib = _native_lib()
    n = ctypes.c_int64(0)
    if lib.emm_pdb_count_atoms(data, ctypes.c_int64(len(data)), ctypes.byref(n)) != 0:
        raise ValueError(lib.emm_pdb_last_error().decode(errors="replace"))
    cap = max(n.value, 1)
    serial = np.zeros(cap, np.int32); resnum = np.zeros(cap, np.int32)
    name = np.zeros(4 * cap, np.uint8); resname = np.zeros(4 * cap, np.uint8); chain = np.zeros(2 * cap, np.uint8)
    segment = np.zeros(4 * cap, np.uint8); element = np.zeros(2 * cap, np.uint8)
    altloc = np.zeros(cap, np.uint8); icode = np.zeros(cap, np.uint8)
    xyz = np.zeros((cap, 3), np.float64); occ = np.zeros(cap, np.float64); bfac = np.zeros(cap, np.float64)
    charge = np.zeros(cap, np.int8)
    header = ctypes.create_string_buffer(5)
    got = ctypes.c_int64(0)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.emm_pdb_parse_ex(data, ctypes.c_int64(len(data)), ctypes.c_int32(flags), ctypes.c_int64(cap), p(serial),
                              p(name), p(altloc), p(resname), p(chain), p(resnum), p(icode), p(xyz), p(occ), p(bfac),
                              p(segment), p(element), p(charge), header, ctypes.byref(got))
    if rc != 0:
        raise ValueError(lib.emm_pdb_last_error().decode(errors="replace"))
    k = got.value
    cols = _columns_from_native(k, serial[:k], name[:4 * k], altloc[:k], resname[:4 * k], chain[:2 * k], resnum[:k],
                                icode[:k], occ[:k], bfac[:k], segment[:4 * k], element[:2 * k], charge[:k])
    return cols, xyz[:k], (header.value.decode() or None)


def _first_cif_token(data: bytes) -> bytes:
    """First token of a text after white space and ``#`` comments (where an mmCIF file has ``data_<name>``)."""
    for line in data[:65536].splitlines():
        line = line.strip()
        if line and not line.startswith(b"#"):
            return line.split()[0]
    return b""


def _looks_like_cif(data: bytes) -> bool:
    return _first_cif_token(data)[:5].lower() == b"data_"


def _cif_block_name(data: bytes) -> Optional[str]:
    return _first_cif_token(data)[5:].decode("ascii", "replace") or None


class Molecule:
    """An ordered atom list (mirror of ``pyjess.Molecule``).

    Columns are NumPy arrays of equal length; ``xyz`` is float64 ``[N, 3]`` holding exactly the
    doubles ``float()`` parses from the PDB text (the oracle and the CUDA path both consume
    these, so "identical inputs" means these arrays).
    """

    __slots__ = ("id", "xyz", "_cols", "_digest", "_atoms")

    def __init__(self, atoms: Sequence[Atom] = (), id: Optional[str] = None):
        self.id = id
        n = len(atoms)
        self._cols = _Columns({
            k: np.asarray([getattr(a, k) for a in atoms], dtype=dt) if n else np.zeros(0, dtype=dt)
            for k, dt in _COLUMNS
        })
        self.xyz = np.asarray([(a.x, a.y, a.z) for a in atoms], dtype=np.float64).reshape(n, 3)
        self._digest = None
        self._atoms = None

    # -- construction -----------------------------------------------------------------------
    @classmethod
    def _from_columns(cls, cols, xyz, id) -> "Molecule":
        self = cls.__new__(cls)
        self.id = id
        self._cols = cols if isinstance(cols, _Columns) else _Columns(cols)
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self._digest = None
        self._atoms = None
        return self

    @classmethod
    def load(cls, file: Union[str, os.PathLike, IO[str]], format: str = "detect", id: Optional[str] = None,
             use_author: bool = False) -> "Molecule":
        """Read a PDB or mmCIF file (path, text or binary file object; gzip-compressed content is
        inflated).  ``format``: ``"detect"`` (by content: a ``data_`` block header means mmCIF), ``"pdb"``
        or ``"cif"``, as ``pyjess.Molecule.load``; ``use_author`` reads the ``auth_*`` identifiers of an
        mmCIF file instead of the ``label_*`` ones.  The reference passes a path and an id only
        (``jess_run.py:538``).  OS errors propagate unchanged (``FileNotFoundError`` /
        ``IsADirectoryError`` -> CLI errno, ``enzymm/_cli.py:318-328``)."""
        if format not in ("detect", "pdb", "cif"):
            raise ValueError(f"invalid value for `format` argument: {format!r}")
        if isinstance(file, (str, os.PathLike)):
            with open(os.fspath(file), "rb") as handle:
                data = handle.read()
        else:
            data = file.read()
            if isinstance(data, str):
                data = data.encode("ascii", "replace")
        if data[:2] == b"\x1f\x8b":
            import gzip
            data = gzip.decompress(data)
        is_cif = _looks_like_cif(data)
        if format != "detect" and is_cif != (format == "cif"):
            raise ValueError(f"the input is not in {format!r} format")
        cols, xyz, header_id = _parse_pdb_native(data, CIF_AUTHOR if use_author else 0)
        if is_cif and id is None:
            header_id = _cif_block_name(data) or header_id        # the whole block name, not only four characters
        return cls._from_columns(cols, xyz, id if id is not None else header_id)

    @classmethod
    def loads(cls, text: str, format: str = "detect", id: Optional[str] = None, use_author: bool = False) -> "Molecule":
        return cls.load(io.StringIO(text), format=format, id=id, use_author=use_author)

    # -- pyjess surface ---------------------------------------------------------------------
    def conserved(self, cutoff: float = 0.0) -> "Molecule":
        """New molecule with the atoms whose ``temperature_factor >= cutoff``."""
        keep = self._cols["temperature_factor"] >= float(cutoff)
        return self.select(keep)

    def select(self, keep: np.ndarray) -> "Molecule":
        return Molecule._from_columns(self._cols.take(keep), self.xyz[keep], self.id)

    def copy(self) -> "Molecule":
        return Molecule._from_columns(self._cols.take(slice(None), copy=True), self.xyz.copy(), self.id)

    def with_xyz(self, xyz: np.ndarray) -> "Molecule":
        """Same atoms, other coordinates (used for ``Hit.molecule(transform=True)``)."""
        return Molecule._from_columns(self._cols, xyz, self.id)

    def column(self, name: str) -> np.ndarray:
        return self._cols[name]

    def __len__(self) -> int:
        return self.xyz.shape[0]

    def __bool__(self) -> bool:
        return self.xyz.shape[0] > 0

    def atom(self, i: int) -> Atom:
        c = self._cols
        x, y, z = self.xyz[i]
        return Atom(
            serial=c["serial"][i], name=c.text_at("name", i), altloc=c.text_at("altloc", i),
            residue_name=c.text_at("residue_name", i), chain_id=c.text_at("chain_id", i),
            residue_number=c["residue_number"][i], insertion_code=c.text_at("insertion_code", i),
            x=x, y=y, z=z, occupancy=c["occupancy"][i],
            temperature_factor=c["temperature_factor"][i], segment=c.text_at("segment", i),
            element=c.text_at("element", i), charge=c["charge"][i],
        )

    def __getitem__(self, i):
        if isinstance(i, slice):
            idx = np.arange(len(self))[i]
            return self.select(idx)
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return self.atom(i)

    def __iter__(self) -> Iterator[Atom]:
        if self._atoms is None:
            self._atoms = [self.atom(i) for i in range(len(self))]
        return iter(self._atoms)

    # hashable + comparable: molecules are dict keys in Matcher.run (jess_run.py:915)
    def _content_digest(self) -> bytes:
        if self._digest is None:
            h = hashlib.blake2b(digest_size=16)
            h.update(repr(self.id).encode())
            h.update(self.xyz.tobytes())
            for k, _ in _COLUMNS:
                h.update(self._cols.canonical_bytes(k))
            self._digest = h.digest()
        return self._digest

    def __eq__(self, other):
        if not isinstance(other, Molecule):
            return NotImplemented
        return self is other or self._content_digest() == other._content_digest()

    def __ne__(self, other):
        if not isinstance(other, Molecule):
            return NotImplemented
        return not self.__eq__(other)

    def __hash__(self):
        # cheap and consistent with __eq__ (equal content => equal id, size and coordinates); the full
        # digest is only computed when two distinct objects with the same hash are compared.  Hashing
        # every column of every molecule cost more than the GPU search in Matcher.run (round 2).
        n = len(self)
        probe = self.xyz[:: max(1, n // 8)][:8].tobytes() if n else b""
        return hash((self.id, n, probe))

    def __repr__(self):
        return f"Molecule(id={self.id!r}, atoms={len(self)})"


def load_many(paths: Sequence[Union[str, os.PathLike]], ids: Optional[Sequence[Optional[str]]] = None,
              threads: int = 0, use_author: bool = False) -> List[Molecule]:
    """Read and parse many PDB / mmCIF files (gzip-compressed or not, told apart by content) on a native
    thread pool (``emm_pdb_load_files_ex``); the returned molecules are views into one SoA batch.
    ``ids`` default to each file's HEADER idCode (mmCIF: the data block name when it has at most four
    characters)."""
    lib = _native_lib()
    paths = [os.fspath(p) for p in paths]
    if not paths:
        return []
    for p in paths:                      # keep Python's exception types for the CLI's errno mapping
        if os.path.isdir(p):
            raise IsADirectoryError(21, "Is a directory", p)
        if not os.path.exists(p):
            raise FileNotFoundError(2, "No such file or directory", p)
    arr = (ctypes.c_char_p * len(paths))(*[p.encode() for p in paths])
    handle = ctypes.c_void_p()
    n_threads = threads or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count() or 1)
    rc = lib.emm_pdb_load_files_ex(arr, ctypes.c_int32(len(paths)), ctypes.c_int32(n_threads),
                                   ctypes.c_int32(CIF_AUTHOR if use_author else 0), ctypes.byref(handle))
    if rc != 0:
        raise ValueError(lib.emm_pdb_last_error().decode(errors="replace"))
    owner = _NativeBatch(lib, handle)       # columns stay views of the native buffers: no copies
    c = _PdbColumns()
    lib.emm_pdb_batch_columns(handle, ctypes.byref(c))
    n, nf = c.n_atoms, c.n_files
    grab = lambda ptr, dtype, count: _native_view(owner, ptr, dtype, count)
    off = grab(c.atom_off, np.int64, nf + 1)
    cols = _columns_from_native(
        n, grab(c.serial, np.int32, n), grab(c.name, np.uint8, 4 * n), grab(c.altloc, np.uint8, n),
        grab(c.resname, np.uint8, 4 * n), grab(c.chain, np.uint8, 2 * n), grab(c.resnum, np.int32, n),
        grab(c.icode, np.uint8, n), grab(c.occupancy, np.float64, n), grab(c.bfactor, np.float64, n),
        grab(c.segment, np.uint8, 4 * n), grab(c.element, np.uint8, 2 * n), grab(c.charge, np.int8, n))
    xyz = grab(c.xyz, np.float64, 3 * n).reshape(n, 3)
    headers = grab(c.header_id, np.uint8, 5 * nf).reshape(nf, 5)
    out = []
    for i in range(nf):
        lo, hi = int(off[i]), int(off[i + 1])
        hid = bytes(headers[i]).split(b"\0")[0].decode() or None
        mol_id = ids[i] if ids is not None and ids[i] is not None else hid
        out.append(Molecule._from_columns(cols.take(slice(lo, hi)), xyz[lo:hi], mol_id))
    return out
