"""Molecules -> ``PackedBatch``: the host half of north_star subsystem (1), batched upload.

Turns ``Molecule`` objects (the ``pyjess.Molecule`` stand-in) into the SoA columns of
``emm_batch``: float64 coordinates exactly as parsed, typing class per atom, residue ordinal per
atom (a residue = one (chain_id, residue_number), SURVEY.md 8c rule 4), B-factor, chain code.
"""
from __future__ import annotations

import ctypes
import os
import warnings
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from .engine import PackedBatch
from .library import CompiledLibrary
from .structures import Molecule, _Columns, _NativeBatch, _native_lib, _native_view

__all__ = ["residue_ordinals", "pack_molecules", "pack_files", "chain_codes", "write_corpus", "read_corpus",
           "Corpus", "slice_batch", "is_corpus", "CORPUS_SUFFIX"]


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
                   with_chain: bool = True, block: int = 1024) -> PackedBatch:
    """Concatenate molecules into one ``PackedBatch``.  Vectorised over blocks of molecules: atom
    kinds (residue name, atom name) are compared as packed integers and classified once per
    distinct kind (``library.class_of``); residue ordinals come from one run-length pass.  A block
    holding a molecule with a split residue goes through the per-molecule path, which reorders."""
    if molecules and all(isinstance(m._cols, _Columns) and {"name", "residue_name", "chain_id"} <= set(m._cols.raw)
                         for m in molecules):
        return _pack_molecules_native(molecules, library, with_chain)
    sizes = np.asarray([len(m) for m in molecules], dtype=np.int64)
    atom_off = np.zeros(len(molecules) + 1, dtype=np.int64)
    np.cumsum(sizes, out=atom_off[1:])
    total = int(atom_off[-1])
    xyz = np.empty((total, 3), dtype=np.float64)
    klass = np.empty(total, dtype=np.uint16)
    residue = np.empty(total, dtype=np.int32)
    bfactor = np.empty(total, dtype=np.float32)
    chain = np.empty(total, dtype=np.uint16)
    atom_id: Optional[np.ndarray] = None
    for b0 in range(0, len(molecules), block):
        mols = molecules[b0:b0 + block]
        lo, hi = int(atom_off[b0]), int(atom_off[b0 + len(mols)])
        if hi == lo:
            continue
        cat = lambda f: np.concatenate([f(m) for m in mols])

        def packed(column: str, width: int) -> np.ndarray:
            """Text column of the block as integers: one view of the concatenated raw bytes when every
            molecule still holds them (native reader), else per-molecule conversion."""
            if all(column in m._cols.raw for m in mols):
                raw = np.ascontiguousarray(np.concatenate([m._cols.raw[column] for m in mols]))
                return raw.view("<u4" if width == 4 else "<u2").reshape(-1)
            return cat(lambda m: m._cols.packed_u32(column))

        res32 = packed("residue_name", 4)
        name32 = packed("name", 4)
        codes = packed("chain_id", 2).astype(np.uint16)
        resnum = cat(lambda m: m.column("residue_number"))
        bsizes = sizes[b0:b0 + len(mols)]
        # residue ordinals: runs of equal (molecule, chain, residue number)
        mol_idx = np.repeat(np.arange(len(mols), dtype=np.int64), bsizes)
        rkey = (mol_idx << 48) | (codes.astype(np.int64) << 32) | (resnum.astype(np.int64) & 0xFFFFFFFF)
        change = np.empty(hi - lo, dtype=bool)
        change[0] = True
        np.not_equal(rkey[1:], rkey[:-1], out=change[1:])
        starts = rkey[change]
        if len(np.unique(starts)) != len(starts):          # some residue is split: reorder per molecule
            atom_id = _pack_block_per_molecule(mols, b0, atom_off, sizes, library, xyz, klass, residue, bfactor, chain,
                                               atom_id, total)
            continue
        runs = np.cumsum(change) - 1
        first = runs[(atom_off[b0:b0 + len(mols)] - lo)[bsizes > 0]]
        residue[lo:hi] = runs - np.repeat(first, bsizes[bsizes > 0])
        # typing classes: one class_of per distinct atom kind of the block
        key = (res32.astype(np.uint64) << np.uint64(32)) | name32.astype(np.uint64)
        klass[lo:hi] = library.classify_keys(key)
        chain[lo:hi] = codes
        bfactor[lo:hi] = cat(lambda m: m.column("temperature_factor"))
        xyz[lo:hi] = cat(lambda m: m.xyz)
        if atom_id is not None:
            atom_id[lo:hi] = np.arange(hi - lo, dtype=np.int64) - np.repeat(atom_off[b0:b0 + len(mols)] - lo, bsizes)
    return PackedBatch(atom_off, xyz, klass, residue, bfactor, chain if with_chain else None, atom_id)


def _pack_molecules_native(molecules: Sequence[Molecule], library: CompiledLibrary, with_chain: bool) -> PackedBatch:
    """Molecules that still hold their raw name bytes (native reader) are packed by
    ``emm_pack_columns`` on the native thread pool: same result as the NumPy path below."""
    lib = _native_lib()
    n = len(molecules)
    keep = []                                  # contiguous arrays the pointers refer to, kept alive for the call

    def pointers(get, dtype):
        out = (ctypes.c_void_p * n)()
        for i, m in enumerate(molecules):
            arr = np.ascontiguousarray(get(m), dtype=dtype)
            keep.append(arr)
            out[i] = arr.ctypes.data
        return out

    sizes = np.asarray([len(m) for m in molecules], dtype=np.int64)
    handle = ctypes.c_void_p()
    n_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    rc = lib.emm_pack_columns(
        ctypes.c_int32(n), sizes.ctypes.data_as(ctypes.c_void_p),
        pointers(lambda m: m._cols.raw["name"], np.uint8), pointers(lambda m: m._cols.raw["residue_name"], np.uint8),
        pointers(lambda m: m._cols.raw["chain_id"], np.uint8), pointers(lambda m: m.column("residue_number"), np.int32),
        pointers(lambda m: m.xyz, np.float64), pointers(lambda m: m.column("temperature_factor"), np.float64),
        ctypes.c_int32(n_threads), ctypes.byref(handle))
    if rc != 0:
        raise RuntimeError(f"emm_pack_columns failed ({rc})")
    return _packed_from_handle(lib, handle, library, with_chain)[0]


def _pack_block_per_molecule(mols, b0, atom_off, sizes, library, xyz, klass, residue, bfactor, chain, atom_id, total):
    """Per-molecule packing (handles split residues by a stable regrouping + ``atom_id``)."""
    for i, m in enumerate(mols, start=b0):
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
    return atom_id


class _PdbPacked(ctypes.Structure):       # struct emm_pdb_packed
    _fields_ = [("n_files", ctypes.c_int32), ("n_atoms", ctypes.c_int64)] + \
               [(k, ctypes.c_void_p) for k in ("atom_off", "xyz", "kind", "residue", "bfactor", "chain", "atom_id", "klass")] + \
               [("n_kinds", ctypes.c_int32), ("kind_names", ctypes.c_void_p), ("header_id", ctypes.c_void_p),
                ("res_off", ctypes.c_void_p), ("res_key", ctypes.c_void_p), ("residue_count", ctypes.c_void_p)]


def pack_files(paths: Sequence[Union[str, os.PathLike]], library: CompiledLibrary, with_chain: bool = True,
               threads: int = 0, use_author: bool = False, on_error: str = "raise") -> Tuple[PackedBatch, List[Optional[str]]]:
    """PDB / mmCIF files (gzip-compressed or not) -> ``PackedBatch`` on the native thread pool
    (``emm_pdb_pack_files_ex``), without building ``Molecule`` objects: the same columns
    ``pack_molecules(load_many(paths), library)`` gives, at parser speed.  Returns ``(batch, header_ids)``;
    what replaces the per-file ``Molecule.load`` of ``jess_run.py:538-548`` when only the hits are wanted.

    ``on_error="raise"`` (default) keeps the reference's behaviour -- the first unreadable or malformed
    file raises, OS errors with their Python types (``_cli.py:318-328`` maps them to exit codes).
    ``on_error="skip"`` is for screening runs over very many files: such a file is reported with a
    warning, stays in the batch as a structure without atoms (it can have no hits) and is listed in
    ``batch.bad_files`` (index -> message); every other file is packed as usual."""
    if on_error not in ("raise", "skip"):
        raise ValueError(f"on_error must be 'raise' or 'skip', not {on_error!r}")
    lib = _native_lib()
    paths = [os.fspath(p) for p in paths]
    skip = on_error == "skip"
    if not skip:
        for p in paths:
            if os.path.isdir(p):
                raise IsADirectoryError(21, "Is a directory", p)
            if not os.path.exists(p):
                raise FileNotFoundError(2, "No such file or directory", p)
    arr = (ctypes.c_char_p * len(paths))(*[p.encode() for p in paths])
    handle = ctypes.c_void_p()
    n_threads = threads or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count() or 1)
    flags = (1 if use_author else 0) | (2 if skip else 0)          # EMM_PDB_CIF_AUTHOR | EMM_PDB_SKIP_BAD
    rc = lib.emm_pdb_pack_files_ex(arr, ctypes.c_int32(len(paths)), ctypes.c_int32(n_threads), ctypes.c_int32(flags),
                                   ctypes.byref(handle))
    if rc != 0:
        raise ValueError(lib.emm_pdb_last_error().decode(errors="replace"))
    bad = {}
    if skip and paths:
        status = np.zeros(len(paths), dtype=np.int32)
        lib.emm_pdb_batch_file_status(handle, status.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(len(status)))
        for i in np.nonzero(status)[0]:
            bad[int(i)] = lib.emm_pdb_batch_file_message(handle, int(i)).decode(errors="replace")
            warnings.warn(f"skipped: {bad[int(i)]}")
    batch, ids = _packed_from_handle(lib, handle, library, with_chain)
    batch.bad_files = bad
    return batch, ids


def _packed_from_handle(lib, handle, library: CompiledLibrary, with_chain: bool):
    """``(PackedBatch, header ids)`` from a native packed batch: classify its atom kinds through
    ``library`` and wrap the columns without copying them."""
    owner = _NativeBatch(lib, handle)       # the big columns stay views of the native buffers
    c = _PdbPacked()
    if lib.emm_pdb_batch_packed(handle, ctypes.byref(c)) != 0:
        raise RuntimeError("emm_pdb_batch_packed failed")
    n, nf = c.n_atoms, c.n_files

    grab = lambda ptr, dtype, count: _native_view(owner, ptr, dtype, count)

    names = grab(c.kind_names, np.uint8, 8 * c.n_kinds).reshape(-1, 8)
    class_of_kind = np.zeros(max(len(names), 1), dtype=np.uint16)
    for i, row in enumerate(names):
        res = bytes(row[:4]).split(b"\0")[0].decode("ascii", "replace")
        name = bytes(row[4:]).split(b"\0")[0].decode("ascii", "replace")
        class_of_kind[i] = library.class_of(res, name)
    if lib.emm_pdb_batch_classify(handle, class_of_kind.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(len(class_of_kind))) != 0:
        raise RuntimeError("emm_pdb_batch_classify failed")
    lib.emm_pdb_batch_packed(handle, ctypes.byref(c))
    atom_off = grab(c.atom_off, np.int64, nf + 1)
    xyz = grab(c.xyz, np.float64, 3 * n).reshape(n, 3)
    klass = grab(c.klass, np.uint16, n)
    residue = grab(c.residue, np.int32, n)
    bfactor = grab(c.bfactor, np.float32, n)
    chain = grab(c.chain, np.uint16, n) if with_chain else None
    atom_id = grab(c.atom_id, np.int32, n) if c.atom_id else None
    headers = grab(c.header_id, np.uint8, 5 * nf).reshape(nf, 5)
    ids = [bytes(h).split(b"\0")[0].decode() or None for h in headers]
    batch = PackedBatch(atom_off, xyz, klass, residue, bfactor, chain, atom_id)
    # what the results table needs about these structures (tsv.TableWriter), as views of the same buffers
    from .tsv import TableColumns
    res_off = grab(c.res_off, np.int64, nf + 1)
    batch.table = TableColumns(atom_off, grab(c.kind, np.uint32, n), names, residue, res_off,
                               grab(c.res_key, np.uint64, int(res_off[-1])), grab(c.residue_count, np.int32, nf),
                               atom_id, owner)
    return batch, ids


# ---- packed corpus files -------------------------------------------------------------------------
# "at > 10^4 structures/s the text parse is the wall" -- a corpus that is screened more
# than once (another template library, other thresholds) is parsed ONCE into the columns the upload
# wants and kept in a file that is mapped, not read: one header, then the raw arrays of
# ``emm_pdb_packed``.  What is stored is independent of any template library: per atom its KIND
# (residue name, atom name), not its typing class; ``read_corpus`` classifies the few hundred kinds
# through the library at hand and expands them with one array lookup.

CORPUS_SUFFIX = ".emmpack"
_CORPUS_MAGIC = b"EMMPACK1"
_CORPUS_ARRAYS = ("atom_off", "xyz", "kind", "residue", "bfactor", "chain", "atom_id", "kind_names", "header_id",
                  "res_off", "res_key", "residue_count", "name_off", "name_blob")


def is_corpus(path) -> bool:
    return os.fspath(path).endswith(CORPUS_SUFFIX)


def _query_ids(paths: Sequence[str]) -> List[str]:
    """File stems, repeats numbered ``_2``, ``_3`` ... (``load_molecules``, ``jess_run.py:523-536``)."""
    import collections
    from pathlib import Path
    seen = collections.defaultdict(int)
    out = []
    for p in paths:
        stem = Path(p).stem
        seen[stem] += 1
        out.append(stem if seen[stem] == 1 else f"{stem}_{seen[stem]}")
    return out


def write_corpus(paths: Sequence[Union[str, os.PathLike]], out: Union[str, os.PathLike], threads: int = 0,
                 use_author: bool = False, on_error: str = "raise", ids: Optional[Sequence[str]] = None) -> int:
    """Parse and pack ``paths`` (PDB / mmCIF, gzip-compressed or not) on the native thread pool and write
    the packed columns to ``out`` (conventionally ``*.emmpack``).  ``ids``: the query id of every
    structure (default: the file stems, as ``load_molecules`` names them).  Returns the number of
    structures written; with ``on_error="skip"`` an unreadable file is written as a structure without
    atoms (``pack_files``)."""
    import json
    if on_error not in ("raise", "skip"):
        raise ValueError(f"on_error must be 'raise' or 'skip', not {on_error!r}")
    lib = _native_lib()
    paths = [os.fspath(p) for p in paths]
    names = list(ids) if ids is not None else _query_ids(paths)
    if len(names) != len(paths):
        raise ValueError("ids must name every path")
    skip = on_error == "skip"
    if not skip:
        for p in paths:
            if os.path.isdir(p):
                raise IsADirectoryError(21, "Is a directory", p)
            if not os.path.exists(p):
                raise FileNotFoundError(2, "No such file or directory", p)
    arr = (ctypes.c_char_p * max(len(paths), 1))(*[p.encode() for p in paths])
    handle = ctypes.c_void_p()
    n_threads = threads or (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count() or 1)
    flags = (1 if use_author else 0) | (2 if skip else 0)
    rc = lib.emm_pdb_pack_files_ex(arr, ctypes.c_int32(len(paths)), ctypes.c_int32(n_threads), ctypes.c_int32(flags),
                                   ctypes.byref(handle))
    if rc != 0:
        raise ValueError(lib.emm_pdb_last_error().decode(errors="replace"))
    owner = _NativeBatch(lib, handle)
    c = _PdbPacked()
    if lib.emm_pdb_batch_packed(handle, ctypes.byref(c)) != 0:
        raise RuntimeError("emm_pdb_batch_packed failed")
    if skip and paths:
        status = np.zeros(len(paths), dtype=np.int32)
        lib.emm_pdb_batch_file_status(handle, status.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(len(status)))
        for i in np.nonzero(status)[0]:
            warnings.warn(f"skipped: {lib.emm_pdb_batch_file_message(handle, int(i)).decode(errors='replace')}")
    n, nf = c.n_atoms, c.n_files
    grab = lambda ptr, dtype, count: _native_view(owner, ptr, dtype, count)
    res_off = grab(c.res_off, np.int64, nf + 1)
    encoded = [s.encode("utf-8") for s in names]
    name_off = np.zeros(nf + 1, dtype=np.int64)
    np.cumsum([len(e) for e in encoded], out=name_off[1:])
    arrays = {
        "atom_off": grab(c.atom_off, np.int64, nf + 1), "xyz": grab(c.xyz, np.float64, 3 * n).reshape(n, 3),
        "kind": grab(c.kind, np.uint32, n), "residue": grab(c.residue, np.int32, n),
        "bfactor": grab(c.bfactor, np.float32, n), "chain": grab(c.chain, np.uint16, n),
        "atom_id": grab(c.atom_id, np.int32, n) if c.atom_id else None,
        "kind_names": grab(c.kind_names, np.uint8, 8 * c.n_kinds).reshape(-1, 8),
        "header_id": grab(c.header_id, np.uint8, 5 * nf).reshape(nf, 5),
        "res_off": res_off, "res_key": grab(c.res_key, np.uint64, int(res_off[-1]) if nf else 0),
        "residue_count": grab(c.residue_count, np.int32, nf),
        "name_off": name_off, "name_blob": np.frombuffer(b"".join(encoded), dtype=np.uint8),
    }
    layout, offset = {}, 0
    for key in _CORPUS_ARRAYS:
        a = arrays[key]
        if a is None:
            continue
        a = arrays[key] = np.ascontiguousarray(a)
        layout[key] = {"dtype": a.dtype.str, "shape": list(a.shape), "offset": offset}
        offset += (a.nbytes + 63) & ~63
    header = json.dumps({"version": 1, "n_structures": int(nf), "n_atoms": int(n), "arrays": layout}).encode()
    data_start = (16 + len(header) + 63) & ~63
    tmp = os.fspath(out) + f".{os.getpid()}.tmp"
    try:
        with open(tmp, "wb") as f:
            f.write(_CORPUS_MAGIC)
            f.write(np.uint64(len(header)).tobytes())
            f.write(header)
            for key, spec in layout.items():
                f.seek(data_start + spec["offset"])
                f.write(arrays[key].tobytes() if arrays[key].nbytes < (1 << 20) else arrays[key].data)
            f.truncate(data_start + offset)
        os.replace(tmp, os.fspath(out))
    finally:
        if os.path.exists(tmp):                      # a write that failed half way leaves nothing behind
            os.unlink(tmp)
    return int(nf)


class Corpus:
    """One mapped corpus file.  ``chunk(lo, hi, library)`` gives the structures ``[lo, hi)`` as a
    ``PackedBatch`` of their own -- views of the mapped columns with offsets rebased, the typing class
    column expanded from the stored kinds for ``library`` -- so a large corpus is handed to the device
    chunk by chunk without ever being classified, copied or even touched as a whole."""

    def __init__(self, path: Union[str, os.PathLike]):
        import mmap
        self.path = path = os.fspath(path)
        with open(path, "rb") as f:
            meta, data_start = self._header(f, path)
            size = os.fstat(f.fileno()).st_size
            mapped = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) if size > data_start else None
        self._mapped = mapped

        def view(key):
            spec = meta["arrays"].get(key)
            if spec is None:
                return None
            count = int(np.prod(spec["shape"])) if spec["shape"] else 1
            if count == 0:
                return np.zeros(spec["shape"], dtype=np.dtype(spec["dtype"]))
            end = data_start + spec["offset"] + count * np.dtype(spec["dtype"]).itemsize
            if mapped is None or end > size:
                raise ValueError(f"{path}: truncated corpus file")
            return np.frombuffer(mapped, dtype=np.dtype(spec["dtype"]), count=count,
                                 offset=data_start + spec["offset"]).reshape(spec["shape"])

        self._a = {key: view(key) for key in _CORPUS_ARRAYS}
        self.n_structures = len(self._a["atom_off"]) - 1
        self._classes = {}                       # id(library) -> (library, class of every stored kind)

    @staticmethod
    def _header(f, path):
        import json
        head = f.read(16)
        if len(head) < 16 or head[:8] != _CORPUS_MAGIC:
            raise ValueError(f"{path} is not a packed corpus (no {_CORPUS_MAGIC.decode()} header)")
        n_header = int(np.frombuffer(head[8:], dtype=np.uint64)[0])
        meta = json.loads(f.read(n_header).decode())
        if meta.get("version") != 1:
            raise ValueError(f"{path}: unsupported corpus version {meta.get('version')!r}")
        return meta, (16 + n_header + 63) & ~63

    @classmethod
    def size_of(cls, path: Union[str, os.PathLike]) -> int:
        """Number of structures in a corpus file (reads its header only)."""
        with open(os.fspath(path), "rb") as f:
            return int(cls._header(f, os.fspath(path))[0]["n_structures"])

    def classes(self, library: CompiledLibrary) -> np.ndarray:
        """Typing class of every stored kind for ``library`` (a few hundred look-ups, kept per library)."""
        hit = self._classes.get(id(library))
        if hit is not None and hit[0] is library:
            return hit[1]
        names = self._a["kind_names"]
        class_of_kind = np.zeros(max(len(names), 1), dtype=np.uint16)
        for i, row in enumerate(names):
            res = bytes(row[:4]).split(b"\0")[0].decode("ascii", "replace")
            name = bytes(row[4:]).split(b"\0")[0].decode("ascii", "replace")
            class_of_kind[i] = library.class_of(res, name)
        self._classes[id(library)] = (library, class_of_kind)
        return class_of_kind

    def ids(self, lo: int = 0, hi: Optional[int] = None) -> List[str]:
        hi = self.n_structures if hi is None else hi
        off = self._a["name_off"]
        blob = self._a["name_blob"][int(off[lo]):int(off[hi])].tobytes()
        base = int(off[lo])
        return [blob[int(off[i]) - base:int(off[i + 1]) - base].decode("utf-8") for i in range(lo, hi)]

    def chunk(self, lo: int, hi: int, library: CompiledLibrary, with_chain: bool = True) -> PackedBatch:
        from .tsv import TableColumns
        a = self._a
        a0, a1 = int(a["atom_off"][lo]), int(a["atom_off"][hi])
        kind = a["kind"][a0:a1]
        klass = self.classes(library)[kind] if a1 > a0 else np.zeros(0, dtype=np.uint16)
        atom_id = None if a["atom_id"] is None else a["atom_id"][a0:a1]
        batch = PackedBatch(a["atom_off"][lo:hi + 1] - a0, a["xyz"][a0:a1], klass, a["residue"][a0:a1], a["bfactor"][a0:a1],
                            a["chain"][a0:a1] if with_chain else None, atom_id)
        r0, r1 = int(a["res_off"][lo]), int(a["res_off"][hi])
        batch.table = TableColumns(batch.atom_off, kind, a["kind_names"], batch.residue, a["res_off"][lo:hi + 1] - r0,
                                   a["res_key"][r0:r1], a["residue_count"][lo:hi], batch.atom_id, self._mapped)
        batch.header_ids = [bytes(h).split(b"\0")[0].decode() or None for h in a["header_id"][lo:hi]]
        batch.bad_files = {}
        return batch


def read_corpus(path: Union[str, os.PathLike], library: CompiledLibrary, with_chain: bool = True
                ) -> Tuple[PackedBatch, List[str]]:
    """A whole corpus file -> ``(PackedBatch, query ids)`` for ``library``: the big columns are views of the
    mapped file (page cache, no parse, no copy); the typing class column is expanded here from the
    stored kinds.  ``batch.table`` carries what the results table needs, ``batch.header_ids`` the
    HEADER idCodes.  (``Corpus`` hands out chunks instead.)"""
    corpus = Corpus(path)
    return corpus.chunk(0, corpus.n_structures, library, with_chain), corpus.ids()


def slice_batch(batch: PackedBatch, lo: int, hi: int) -> PackedBatch:
    """Structures ``[lo, hi)`` of a packed batch as a batch of their own (views, offsets rebased)."""
    a0, a1 = int(batch.atom_off[lo]), int(batch.atom_off[hi])
    cut = lambda col: None if col is None else col[a0:a1]
    out = PackedBatch(batch.atom_off[lo:hi + 1] - a0, batch.xyz[a0:a1], batch.klass[a0:a1], batch.residue[a0:a1],
                      cut(batch.bfactor), cut(batch.chain), cut(batch.atom_id))
    table = getattr(batch, "table", None)
    if table is not None:
        from .tsv import TableColumns
        r0, r1 = int(table.res_off[lo]), int(table.res_off[hi])
        out.table = TableColumns(out.atom_off, table.kind[a0:a1], table.kind_names, out.residue,
                                 table.res_off[lo:hi + 1] - r0, table.res_key[r0:r1], table.residue_count[lo:hi],
                                 out.atom_id, table._owner)
    return out
