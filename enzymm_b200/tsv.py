"""The results table (``Match.dump``, reference ``enzymm/jess_run.py:185-284``) for whole batches of
hit records, without a ``Match`` / ``Hit`` / ``Atom`` object per row.

``Matcher.run`` followed by ``Match.dump`` per match is the reference's way and stays available;
at screening rates (> 10^4 structures/s, several rows per structure) it is the Python objects that
cost the time, not the search.  ``TableWriter`` keeps the reference's semantics --

* rows = the raw best hits of every (molecule, size group) that survive the filter (or all of them
  with ``filter_matches=False``), ``jess_run.py:867-894``;
* ``completeness`` decided per (molecule, size group) BEFORE filtering, ``jess_run.py:738-783``;
* molecules in the order in which they first receive a surviving match, size-major; matches of a
  molecule size-major, then template order; ``match_index`` 1-based per molecule (``_cli.py:272-275``);
* a hit filtered at a distance without logistic models is a ``KeyError`` (``jess_run.py:339-342``)

-- as NumPy operations over the ``emm_hit`` columns, gathers the residue name / chain / residue
number of the matched atoms from the packed batch, and hands the rows to the native formatter
(``emm_tsv_format``, ``csrc/emm_tsv.cpp``), which prints them byte for byte as Python would.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np

from .engine import HIT_NO_MODEL, HIT_PASS, load_cdll
from .library import MAX_TEMPLATE_ATOMS
from .templates import AnnotatedTemplate

__all__ = ["TableColumns", "RowSelector", "TableWriter", "TSV_HEADER"]

TSV_HEADER = [
    "query_id", "pairwise_distance", "match_index", "template_pdb_id", "template_pdb_chains",
    "template_cluster_id", "template_cluster_member", "template_cluster_size",
    "template_effective_size", "template_dimension", "template_mcsa_id", "template_uniprot_id",
    "template_ec", "template_cath", "template_multimeric", "query_multimeric", "query_atom_count",
    "query_residue_count", "rmsd", "log_evalue", "orientation", "preserved_order", "completeness",
    "predicted_correct", "matched_residues", "number_of_mutated_residues",
    "number_of_side_chain_residues_(template,reference)",
    "number_of_metal_ligands_(template,reference)", "number_of_ptm_residues_(template, reference)",
    "total_reference_residues",
]


class TableColumns:
    """What a table row needs about the query structures of one packed batch, besides the hits:
    per atom its kind (residue name, atom name) and residue ordinal, per residue ordinal the chain id
    and residue number, per structure ``Match.query_residue_count``.  Produced by the native packers
    (``emm_pdb_packed``: ``kind``, ``res_off`` / ``res_key``, ``residue_count``)."""

    def __init__(self, atom_off, kind, kind_names, residue, res_off, res_key, residue_count, atom_id=None, owner=None):
        self.atom_off = atom_off
        self.kind = kind
        self.kind_names = np.ascontiguousarray(kind_names).reshape(-1, 8)
        self.residue = residue
        self.res_off = res_off
        self.res_key = res_key
        self.residue_count = residue_count
        self.atom_id = atom_id
        self._owner = owner              # the native batch the arrays are views of
        self._inverse = None

    def packed_position(self, structure: np.ndarray, atoms: np.ndarray) -> np.ndarray:
        """Global row of the per-atom columns for (structure, reported atom index) pairs.  Hits report
        positions in the input file; only files with a split residue were reordered (``atom_id``)."""
        base = self.atom_off[structure]
        if self.atom_id is None:
            return base + atoms
        if self._inverse is None:
            sizes = np.diff(self.atom_off)
            start = np.repeat(self.atom_off[:-1], sizes)
            inverse = np.empty(len(self.atom_id), dtype=np.int64)
            inverse[start + self.atom_id] = np.arange(len(self.atom_id), dtype=np.int64) - start
            self._inverse = inverse
        return base + self._inverse[base + atoms]


class _Rows(ctypes.Structure):        # struct emm_tsv_rows
    _fields_ = [("n_rows", ctypes.c_int64)] + [(k, ctypes.c_void_p) for k in (
        "n_atoms", "resname4", "chain2", "resnum", "rmsd", "log_evalue", "orientation", "match_index", "complete",
        "predicted", "template_index", "structure", "query_id", "query_atom_count", "query_residue_count",
        "tpl_distance", "tpl_static", "tpl_multimeric", "tpl_order_off", "tpl_order", "tpl_annotation")]


def _c_strings(values: Sequence[str]):
    encoded = [v.encode("utf-8") for v in values]
    return (ctypes.c_char_p * max(len(encoded), 1))(*encoded), encoded


class RowSelector:
    """Which hit records of a batch survive as matches, in which order, with which completeness flag:
    ``Matcher.run``'s bookkeeping (``jess_run.py:738-783, 845-894``) as array operations."""

    def __init__(self, matcher, predict_correctness: bool = True):
        self.matcher = matcher
        matcher._compile()
        self.predict_correctness = predict_correctness
        self.bounds = np.asarray([hi for _, _, hi in matcher._groups])
        ident, member, size = matcher._template_identity()
        self.ident = np.asarray(ident, dtype=np.int64)
        self.member = np.asarray(member, dtype=np.int64)
        self.size = np.asarray(size, dtype=np.int64)
        self.n_ident = int(self.ident.max()) + 2 if len(ident) else 1

    def select(self, records: np.ndarray):
        """``(rows, match_index, complete, predicted)``: indices into ``records`` in output order,
        ``Match.index``, ``Match.complete`` and the ``predicted_correct`` column (0 / 1 / 2 = empty)."""
        n = len(records)
        empty = np.zeros(0, dtype=np.int64)
        if n == 0:
            return empty, empty.astype(np.int32), empty.astype(np.uint8), empty.astype(np.uint8)
        tidx = records["template_index"].astype(np.int64)
        sidx = records["structure"].astype(np.int64)
        flags = records["flags"]
        group_of = np.searchsorted(self.bounds, tidx, side="right").astype(np.int64)
        n_groups = len(self.bounds)
        # completeness per (structure, size group, cluster identity), before any filtering
        ident = self.ident[tidx]
        complete = ident < 0
        has = np.nonzero(~complete)[0]
        if len(has):
            key = (sidx[has] * n_groups + group_of[has]) * self.n_ident + ident[has]
            uniq, first, inverse, count = np.unique(key, return_index=True, return_inverse=True, return_counts=True)
            want = self.size[tidx[has]][first]                  # cluster size of the first match of the group
            member = self.member[tidx[has]]
            stride = int(max(want.max(), member.max())) + 2
            in_range = (member >= 1) & (member <= want[inverse])
            distinct = np.unique(inverse[in_range] * stride + member[in_range]) // stride
            n_distinct = np.bincount(distinct, minlength=len(uniq))
            complete[has] = ((count == want) & (n_distinct == want))[inverse]
        passing = (flags & HIT_PASS) != 0
        if self.matcher.filter_matches or self.predict_correctness:
            if ((flags & HIT_NO_MODEL) != 0).any():
                bad = records[(flags & HIT_NO_MODEL) != 0][0]
                distance = self.matcher._get_jess_parameters(self.matcher._ordered[int(bad["template_index"])].effective_size)[1]
                raise KeyError("Missing appropriate model parameters to predict correctness. Encountered either unexpected "
                               f"dictionary structure or no models for the pairwise distance {distance} were provided")
        keep = np.nonzero(passing)[0] if self.matcher.filter_matches else np.arange(n)
        if len(keep) == 0:
            return empty, empty.astype(np.int32), empty.astype(np.uint8), empty.astype(np.uint8)
        # molecules in the order of their first surviving match, size-major (jess_run.py:867-894)
        n_struct = int(sidx.max()) + 1
        first_group = np.full(n_struct, n_groups, dtype=np.int64)
        np.minimum.at(first_group, sidx[keep], group_of[keep])
        order = np.lexsort((tidx[keep], sidx[keep], first_group[sidx[keep]]))
        rows = keep[order]
        srow = sidx[rows]
        new_run = np.r_[True, srow[1:] != srow[:-1]]
        run_start = np.maximum.accumulate(np.where(new_run, np.arange(len(rows)), 0))
        match_index = (np.arange(len(rows)) - run_start + 1).astype(np.int32)
        predicted = passing[rows].astype(np.uint8) if self.predict_correctness else np.full(len(rows), 2, dtype=np.uint8)
        return rows, match_index, complete[rows].astype(np.uint8), predicted


class TableWriter(RowSelector):
    """Rows of the results table for the hit records of a ``Matcher``'s batches."""

    def __init__(self, matcher, predict_correctness: bool = True):
        super().__init__(matcher, predict_correctness)
        templates = matcher._ordered
        distance, static, annotation, order_off, order = [], [], [], [0], []
        for gsize, lo, hi in matcher._groups:
            d = str(matcher._get_jess_parameters(gsize)[1])
            distance.extend([d] * (hi - lo))
        for t in templates:
            c = t.cluster
            static.append("\t".join([
                str(t.pdb_id if t.pdb_id else ""), ",".join(set(r.chain_id for r in t.residues)),
                str(c.id if c else ""), str(c.member if c else ""), str(c.size if c else ""),
                str(t.effective_size), str(t.dimension), str(t.mcsa_id if t.mcsa_id else ""),
                str(t.uniprot_id if t.uniprot_id else ""), ",".join(t.ec if t.ec is not None else ""),
                ",".join(t.cath if t.cath else "")]))
            if isinstance(t, AnnotatedTemplate) and hasattr(t, "number_of_mutated_residues"):
                annotation.append("\t".join([
                    str(t.number_of_mutated_residues), ",".join(str(i) for i in t.number_of_side_chain_residues),
                    ",".join(str(i) for i in t.number_of_metal_ligands),
                    ",".join(str(i) for i in t.number_of_ptm_residues), str(t.total_reference_residues)]))
            else:
                annotation.append("\t" * 5)                       # six empty columns, as the reference writes
            order.extend(t.relative_order)
            order_off.append(len(order))
        self._distance, self._keep1 = _c_strings(distance)
        self._static, self._keep2 = _c_strings(static)
        self._annotation, self._keep3 = _c_strings(annotation)
        self._multimeric = np.asarray([1 if t.multimeric else 0 for t in templates], dtype=np.uint8)
        self._order_off = np.asarray(order_off, dtype=np.int32)
        self._order = np.asarray(order if order else [0], dtype=np.int32)
        self._lib = load_cdll()
        self._lib.emm_tsv_format.restype = ctypes.c_int
        self._lib.emm_tsv_free.restype = None
        self._lib.emm_tsv_free.argtypes = [ctypes.c_void_p]

    @staticmethod
    def header() -> str:
        return "\t".join(TSV_HEADER) + "\n"

    # ---- rows -> text ------------------------------------------------------------------------------------
    def format(self, records: np.ndarray, table: TableColumns, query_ids: Sequence[Optional[str]],
               selection=None) -> bytes:
        """The table rows (no header) of one batch's hit records as UTF-8 bytes."""
        rows, match_index, complete, predicted = selection if selection is not None else self.select(records)
        n_rows = len(rows)
        if n_rows == 0:
            return b""
        rec = records[rows]
        n_atoms = rec["n_atoms"].astype(np.int32)
        structure = rec["structure"].astype(np.int32)
        atoms = np.where(np.arange(MAX_TEMPLATE_ATOMS)[None, :] < n_atoms[:, None], rec["atoms"], 0).astype(np.int64)
        pos = table.packed_position(structure.astype(np.int64)[:, None], atoms)
        names = table.kind_names[table.kind[pos]]                                   # [rows, 32, 8]
        resname4 = np.ascontiguousarray(names[:, :, :4])
        key = table.res_key[table.res_off[structure.astype(np.int64)][:, None] + table.residue[pos]]
        chain = (key >> np.uint64(32)).astype(np.uint16)
        chain2 = np.ascontiguousarray(np.stack([chain & 0xFF, chain >> 8], axis=2).astype(np.uint8))
        resnum = np.ascontiguousarray((key & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32))
        ids, keep_ids = _c_strings([str(q) for q in query_ids])
        atom_count = np.diff(table.atom_off).astype(np.int32)
        residue_count = np.ascontiguousarray(table.residue_count, dtype=np.int32)
        cols = dict(
            n_atoms=n_atoms, resname4=resname4, chain2=chain2, resnum=resnum,
            rmsd=np.ascontiguousarray(rec["rmsd"], dtype=np.float64),
            log_evalue=np.full(n_rows, np.nan), orientation=np.ascontiguousarray(rec["orientation"], dtype=np.float64),
            match_index=np.ascontiguousarray(match_index, dtype=np.int32), complete=np.ascontiguousarray(complete, dtype=np.uint8),
            predicted=np.ascontiguousarray(predicted, dtype=np.uint8),
            template_index=np.ascontiguousarray(rec["template_index"], dtype=np.int32), structure=structure,
            query_atom_count=atom_count, query_residue_count=residue_count, tpl_multimeric=self._multimeric,
            tpl_order_off=self._order_off, tpl_order=self._order)
        ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        arg = _Rows(n_rows, *[None] * 21)
        for name, arr in cols.items():
            setattr(arg, name, ptr(arr))
        arg.query_id = ctypes.cast(ids, ctypes.c_void_p)
        arg.tpl_distance = ctypes.cast(self._distance, ctypes.c_void_p)
        arg.tpl_static = ctypes.cast(self._static, ctypes.c_void_p)
        arg.tpl_annotation = ctypes.cast(self._annotation, ctypes.c_void_p)
        text, length = ctypes.c_void_p(), ctypes.c_int64(0)
        rc = self._lib.emm_tsv_format(ctypes.byref(arg), ctypes.byref(text), ctypes.byref(length))
        if rc != 0:
            raise RuntimeError(f"emm_tsv_format failed ({rc})")
        try:
            return ctypes.string_at(text, length.value)
        finally:
            self._lib.emm_tsv_free(text)
            del keep_ids                         # the encoded ids had to outlive the call

