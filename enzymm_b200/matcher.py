"""``Match`` / ``Matcher`` / ``load_molecules``: EnzyMM's orchestration layer over the CUDA engine.

Same Python-facing API and semantics as ``enzymm/jess_run.py``, but
``Matcher.run`` hands the whole (molecules x templates) problem to the GPU in one batch instead
of one ``Jess(templates).query(molecule)`` per (molecule, size group) on a thread pool
(``jess_run.py:896-988``).  Preserved on purpose (SURVEY.md 5 "quirks"):

* size groups are processed in descending effective size with per-size thresholds
  (``jess_run.py:564-571, 724-736, 930-947``); sizes below 3 only with ``match_small_templates``;
* ``_check_completeness`` runs on the raw best hits of one (molecule, size group) BEFORE
  filtering (``jess_run.py:738-783, 863``);
* filtering keeps ``Match.predicted_correct`` matches; the result dict is keyed in the order
  molecules first receive a surviving match, size-major (``jess_run.py:867-894``);
* ``skip_smaller_hits`` skips a molecule once it holds a surviving match (``jess_run.py:951-958``);
* the majority vote is ">= round(5/2) == 2 of 5" and a missing distance key is a ``KeyError``
  (``jess_run.py:298-346``).

The filter verdict used for ``filter_matches`` / ``skip_smaller_hits`` comes from the fused GPU
filter (``Hit.device_pass``); ``Match.predicted_correct`` is the reference formula in Python and is
what the parity tests compare it with.
"""
from __future__ import annotations

import collections
import csv
import io
import math
import os
import sys
import warnings
from dataclasses import dataclass, field
from functools import cached_property
from pathlib import Path
from typing import ClassVar, Dict, IO, List, Optional, Sequence, Tuple

import numpy as np

from . import pyjess_api as pyjess
from .engine import Engine, EngineError
from .library import CompiledLibrary, load_lr_models
from .packing import pack_molecules
from .pyjess_api import Hit
from .structures import Molecule, load_many
from .templates import AnnotatedTemplate, Template, Vec3, check_template, rank_order

__all__ = ["LogisticRegressionModel", "Match", "Matcher", "load_molecules"]

PROTEINOGENIC_AMINO_ACIDS = ["ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE",
                             "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL"]
SPECIAL_AMINO_ACIDS = ["ASX", "GLX", "SEC", "PYL", "UNK", "MSE", "SEP", "TPO", "PTR", "HYP", "CME",
                       "CSO", "CSD", "PCA", "MLY", "DAL", "DAR", "DSG", "ORN", "PTM"]
_COUNTED_RESIDUES = frozenset(PROTEINOGENIC_AMINO_ACIDS + SPECIAL_AMINO_ACIDS)

_ONE_LETTER_ELEMENTS = frozenset("HBCNOFPSKVYIWU")

_TSV_HEADER = [
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


@dataclass(frozen=True)
class LogisticRegressionModel:
    """f(x) = 1 / (1 + e^-(b0 + b1*rmsd + b2*orientation)) with its decision threshold
    (``jess_run.py:39-56``; the attribute keeps upstream's spelling)."""

    coefficents: List[float]
    intercept: float
    threshold: float


def _pdb_atom_line(atom) -> str:
    """Fixed-column ATOM record as the reference writes it (``jess_run.py:125-145``): one-letter
    elements shift the atom name one column right."""
    altloc = atom.altloc if atom.altloc is not None else ""
    name = f"  {atom.name:<3s}" if atom.element in _ONE_LETTER_ELEMENTS else f" {atom.name:<4s}"
    return (f"ATOM  {atom.serial:>5}{name}{altloc:<1}{atom.residue_name:<3}{atom.chain_id:>2}"
            f"{atom.residue_number:>4}{atom.insertion_code:1s}   {atom.x:>8.3f}{atom.y:>8.3f}{atom.z:>8.3f}"
            f"{atom.occupancy:>6.2f}{atom.temperature_factor:>6.2f}      {atom.segment:<4s}{atom.element:>2s} \n")


@dataclass
class Match:
    """A ``Hit`` plus EnzyMM's derived quantities (``jess_run.py:59-496``)."""

    hit: Hit
    complete: bool = field(default=False)
    pairwise_distance: float = field(default=0)
    index: int = field(default=0)
    _logistic_regression_models: ClassVar[Dict[str, Dict[str, List[LogisticRegressionModel]]]] = {}

    # ---- geometry ------------------------------------------------------------------------------
    @cached_property
    def atom_triplets(self):
        """Matched atoms (template frame) grouped three by three; every triplet must come from
        one query residue (``jess_run.py:348-373``)."""
        atoms = self.hit.atoms(transform=True)
        triplets = []
        for start in range(0, len(atoms), 3):
            triplet = tuple(atoms[start:start + 3])
            if len(triplet) != 3:
                raise ValueError(f"Failed to construct residues. Got only {len(triplet)} ATOM lines")
            owners = {(a.residue_name, a.chain_id, a.residue_number) for a in triplet}
            if len(owners) != 1:
                raise ValueError(f"Mixed up atom triplets {owners}. The atoms come from different residues!")
            triplets.append(triplet)
        return triplets

    @property
    def matched_residues(self) -> List[Tuple[str, str, str]]:
        return [(t[0].residue_name, t[0].chain_id, str(t[0].residue_number)) for t in self.atom_triplets]

    @property
    def multimeric(self) -> bool:
        atoms = self.hit.atoms()
        return any(a.chain_id != atoms[0].chain_id for a in atoms)

    @property
    def preserved_resid_order(self) -> bool:
        if self.hit.template.multimeric or self.multimeric:
            return False
        return rank_order([t[0].residue_number for t in self.atom_triplets]) == self.hit.template.relative_order

    @cached_property
    def match_vector_list(self) -> List[Vec3]:
        """Orientation vector of every matched residue, in the template frame (``jess_run.py:425-452``)."""
        vectors = []
        for triplet, residue in zip(self.atom_triplets, self.hit.template.residues):
            first, second = residue.orientation_vector_indices
            if second == 9:
                centre = triplet[first]
                side_a, side_b = [a for a in triplet if a != centre]
                vectors.append((Vec3.from_xyz(side_a) + Vec3.from_xyz(side_b)) / 2 - Vec3.from_xyz(centre))
            else:
                vectors.append(Vec3.from_xyz(triplet[second]) - Vec3.from_xyz(triplet[first]))
        return vectors

    @property
    def template_vector_list(self) -> List[Vec3]:
        return [r.orientation_vector for r in self.hit.template.residues]

    @property
    def orientation(self) -> float:
        """Mean angle (radians) between template and query residue orientations (``jess_run.py:461-478``)."""
        tv, mv = self.template_vector_list, self.match_vector_list
        if len(tv) != len(mv):
            raise ValueError("Vector lists for Template and matching Query structure had different lengths.")
        angles = [t.angle_to(m) for t, m in zip(tv, mv)]
        return sum(angles) / len(angles)

    # ---- logistic filter --------------------------------------------------------------------------
    @property
    def predicted_correct(self) -> bool:
        """Majority vote of the logistic models for (effective size, pairwise distance); sizes
        without models pass (``jess_run.py:298-346``)."""
        models_by_size = self._logistic_regression_models
        size_key = str(self.hit.template.effective_size)
        if size_key not in models_by_size:
            return True
        try:
            models = models_by_size[size_key][str(self.pairwise_distance)]
            votes = 0
            for model in models:
                z = model.intercept + model.coefficents[0] * self.hit.rmsd + model.coefficents[1] * self.orientation
                votes += (1 / (1 + math.e ** -z)) >= model.threshold
            return bool(votes >= round(len(models) / 2, 0))
        except KeyError as exc:
            raise KeyError(
                "Missing appropriate model parameters to predict correctness. Encountered either unexpected "
                f"dictionary structure or no models for the pairwise distance {self.pairwise_distance} were provided"
            ) from exc
        except IndexError as exc:
            raise IndexError("Missing coefficients for both RMSD and Residue Orientation. "
                             "Expecting models with 2 coeficients.") from exc

    def get_identifying_attributes(self) -> Tuple[int, int, int]:
        t = self.hit.template
        return (t.mcsa_id, t.cluster.id, t.dimension)

    # ---- query statistics ---------------------------------------------------------------------------
    @property
    def query_atom_count(self) -> int:
        return len(self.hit.molecule())

    @property
    def query_residue_count(self) -> int:
        """Distinct residue NUMBERS among amino-acid residues, chain ignored (``jess_run.py:487-496``)."""
        mol = self.hit.molecule()
        names = mol.column("residue_name")
        keep = np.isin(names, list(_COUNTED_RESIDUES))
        return int(len(np.unique(mol.column("residue_number")[keep])))

    # ---- writers --------------------------------------------------------------------------------------
    def dumps(self, header: bool = False) -> str:
        buffer = io.StringIO()
        self.dump(buffer, header=header)
        return buffer.getvalue()

    def dump(self, file: IO[str], header: bool = False, predict_correctness: bool = True):
        """One TSV row (``jess_run.py:185-284``)."""
        writer = csv.writer(file, dialect="excel-tab", delimiter="\t", lineterminator="\n")
        if header:
            writer.writerow(_TSV_HEADER)
        t = self.hit.template
        c = t.cluster
        row = [
            str(self.hit.molecule().id), str(self.pairwise_distance), str(self.index),
            str(t.pdb_id if t.pdb_id else ""), ",".join(set(r.chain_id for r in t.residues)),
            str(c.id if c else ""), str(c.member if c else ""), str(c.size if c else ""),
            str(t.effective_size), str(t.dimension), str(t.mcsa_id if t.mcsa_id else ""),
            str(t.uniprot_id if t.uniprot_id else ""), ",".join(t.ec if t.ec is not None else ""),
            ",".join(t.cath if t.cath else ""), str(t.multimeric), str(self.multimeric),
            str(self.query_atom_count), str(self.query_residue_count), str(round(self.hit.rmsd, 5)),
            str(round(self.hit.log_evalue, 5)), str(round(self.orientation, 5)),
            str(self.preserved_resid_order), str(self.complete),
            str(self.predicted_correct) if predict_correctness else "",
            ",".join("_".join(r) for r in self.matched_residues),
        ]
        if isinstance(t, AnnotatedTemplate) and hasattr(t, "number_of_mutated_residues"):
            row.extend([
                str(t.number_of_mutated_residues), ",".join(str(i) for i in t.number_of_side_chain_residues),
                ",".join(str(i) for i in t.number_of_metal_ligands),
                ",".join(str(i) for i in t.number_of_ptm_residues), str(t.total_reference_residues),
            ])
        else:
            row.extend(["", "", "", "", "", ""])
        writer.writerow(row)

    def dump2pdb(self, file: IO[str], include_query: bool = False, transform: bool = False):
        """Matched atoms (optionally preceded by the whole query) as PDB text (``jess_run.py:96-183``)."""
        mol_id = self.hit.molecule().id
        if include_query:
            file.write(f"HEADER MOLECULE_ID {mol_id}\n")
            for atom in self.hit.molecule(transform=transform):
                file.write(_pdb_atom_line(atom))
            file.write("END\n\n")
        t = self.hit.template
        file.write(f"HEADER {self.predicted_correct} MATCH {mol_id} {self.index}\n")
        file.write(f'REMARK TEMPLATE_PDB {t.pdb_id}_{",".join(set(r.chain_id for r in t.residues))}\n')
        if t.cluster:
            file.write(f"REMARK TEMPLATE CLUSTER {t.cluster.id}_{t.cluster.member}_{t.cluster.size}\n")
        if t.represented_sites:
            file.write(f"REMARK TEMPLATE RESIDUES {t.template_id_string}\n")
        file.write(f"REMARK MOLECULE_ID {mol_id}\n")
        file.write(f"REMARK MATCH INDEX {self.index}\n")
        file.write("REMARK TEMPLATE COORDINATE FRAME\n" if transform else "REMARK QUERY COORDINATE FRAME\n")
        for atom in self.hit.atoms(transform=transform):
            file.write(_pdb_atom_line(atom))
        file.write("END\n\n")


def _install_lr_models():
    Match._logistic_regression_models = {
        size: {dist: [LogisticRegressionModel([c0, c1], b0, thr) for c0, c1, b0, thr in models]
               for dist, models in by_dist.items()}
        for size, by_dist in load_lr_models().items()
    }


_install_lr_models()


def load_molecules(molecule_paths: Sequence[Path], conservation_cutoff: float = 0, warn: bool = False,
                   use_author: bool = False) -> List[Molecule]:
    """Load query structures; repeated file stems get ``_2``, ``_3`` ... ids (``jess_run.py:523-556``).
    PDB and mmCIF files, gzip-compressed or not, are told apart by content (``Molecule.load``'s
    ``format="detect"``); ``use_author`` picks the ``auth_*`` identifiers of mmCIF files.

    As upstream, ``conserved()`` is called and its RESULT DISCARDED (``jess_run.py:541-542``), so
    the cutoff does not mask anything on this path.  The intended masking
    is available through ``Matcher(conservation_cutoff=..., apply_conservation_mask=True)``.
    """
    molecules: List[Molecule] = []
    seen: Dict[str, int] = collections.defaultdict(int)
    ids = []
    for path in molecule_paths:
        stem = Path(path).stem
        seen[stem] += 1
        ids.append(stem if seen[stem] == 1 else f"{stem}_{seen[stem]}")
    # native, multi-threaded ingest (emm_pdb_load_files); OS errors keep their Python types
    for path, mol in zip(molecule_paths, load_many([str(p) for p in molecule_paths], ids=ids, use_author=use_author)):
        if conservation_cutoff:
            mol.conserved(conservation_cutoff)
        if mol:
            molecules.append(mol)
        elif warn:
            warnings.warn(f"received an empty molecule from {path}")
    if not molecules and warn:
        warnings.warn("received no molecules from input")
    return molecules


def _available_cpus() -> int:
    return len(os.sched_getaffinity(0)) if sys.platform == "linux" else (os.cpu_count() or 1)


class Matcher:
    """Match a list of templates against query molecules on the GPU (``jess_run.py:559-1000``)."""

    _DEFAULT_JESS_PARAMS = {
        3: {"rmsd": 2, "distance": 0.9, "max_dynamic_distance": 0.9},
        4: {"rmsd": 2, "distance": 1.7, "max_dynamic_distance": 1.7},
        5: {"rmsd": 2, "distance": 2.0, "max_dynamic_distance": 2.0},
        6: {"rmsd": 2, "distance": 2.0, "max_dynamic_distance": 2.0},
        7: {"rmsd": 2, "distance": 2.0, "max_dynamic_distance": 2.0},
        8: {"rmsd": 2, "distance": 2.0, "max_dynamic_distance": 2.0},
    }

    def __init__(self, templates: List[Template], jess_params: Optional[Dict[int, Dict[str, float]]] = None,
                 conservation_cutoff: int = 0, warn: bool = False, verbose: bool = False,
                 skip_smaller_hits: bool = False, match_small_templates: bool = False,
                 cpus: Optional[int] = None, filter_matches: bool = True, console=None,
                 *, device: int = 0, max_candidates: int = 10000, apply_conservation_mask: bool = False):
        self.templates = templates
        self.cpus = _available_cpus() if cpus is None else cpus
        self.conservation_cutoff = conservation_cutoff
        self.warn = warn
        self.verbose = verbose
        self.skip_smaller_hits = skip_smaller_hits
        self.match_small_templates = match_small_templates
        self.filter_matches = filter_matches
        self.jess_params = self._DEFAULT_JESS_PARAMS if jess_params is None else jess_params
        self.console = console
        self.device = device
        self.max_candidates = max_candidates
        self.apply_conservation_mask = apply_conservation_mask

        if len(set(self.templates)) < len(self.templates):
            raise ValueError("Duplicate templates were found.")
        if self.cpus <= 0:
            self.cpus = max(1, _available_cpus() + self.cpus)

        self.verbose_print(f"PyJess Version: {pyjess.__version__}")
        self.verbose_print(f"Running on {self.cpus} Thread(s)")
        self.verbose_print(f"Warnings are set to {self.warn}")
        self.verbose_print(f"Skip_smaller_hits search is set to {self.skip_smaller_hits}")
        if self.conservation_cutoff:
            self.verbose_print(f"Conservation Cutoff set to {self.conservation_cutoff}")

        self.templates_by_effective_size: Dict[int, List[Template]] = collections.defaultdict(list)
        for template in templates:
            if check_template(template, warn=self.warn):
                self.templates_by_effective_size[template.effective_size].append(template)
        if self.verbose:
            shown = {s: len(v) for s, v in self.templates_by_effective_size.items()
                     if self.match_small_templates or s >= 3}
            print(f"Templates by effective size: {collections.OrderedDict(sorted(shown.items()))}")
        self.template_effective_sizes = sorted(self.templates_by_effective_size, reverse=True)

        if self.warn:
            small = [t for s in self.template_effective_sizes if s < 3 for t in self.templates_by_effective_size[s]]
            if small:
                tail = ("For small templates Jess parameters for templates of 3 residues will be used."
                        if self.match_small_templates else
                        "These will be excluded since these templates are too general.")
                warnings.warn(f"{len(small)} Templates with an effective size smaller than 3 defined "
                              f"sidechain residues were supplied.\n{tail}")
                self.verbose_print("The templates with the following ids are too small:")
                self.verbose_print([t.id for t in small])
        self._engine: Optional[Engine] = None
        self._compiled: Optional[CompiledLibrary] = None
        self._groups: List[Tuple[int, int, int]] = []
        self._scan_lanes = None
        self.hits_per_structure = 64          # initial hit-buffer sizing; grown on demand
        self.skipped_structures: Dict[int, int] = {}   # batch index -> status of structures the device refused
        self._hit_floor = 1024

    def verbose_print(self, *args):
        if self.verbose:
            print(*args)

    def _get_jess_parameters(self, template_size: int) -> Tuple[float, float, float]:
        p = self.jess_params[min(max(template_size, 3), 8)]
        return p["rmsd"], p["distance"], p["max_dynamic_distance"]

    @staticmethod
    def _check_completeness(matches: List[Match]) -> List[Match]:
        """A match is complete when every member of its template cluster hit the same molecule
        in this size group; matches without cluster / M-CSA id are complete (``jess_run.py:738-783``)."""
        grouped: Dict[tuple, List[Match]] = collections.defaultdict(list)
        for match in matches:
            t = match.hit.template
            if t.mcsa_id is not None and t.cluster is not None:
                grouped[match.get_identifying_attributes()].append(match)
            else:
                match.complete = True
        for members in grouped.values():
            want = list(range(1, members[0].hit.template.cluster.size + 1))
            have = sorted(m.hit.template.cluster.member for m in members)
            if have == want:
                for m in members:
                    m.complete = True
        return matches

    # ---- engine ---------------------------------------------------------------------------------------
    def _active_sizes(self) -> List[int]:
        return [s for s in self.template_effective_sizes if s >= 3 or self.match_small_templates]

    def _compile(self) -> CompiledLibrary:
        """Every searched size group in ONE compiled library (size-descending, caller order inside a
        group) with per-template thresholds.  Host work only: no device is touched."""
        if self._compiled is None:
            ordered: List[Template] = []
            rmsd, dist, dyn = [], [], []
            self._groups = []
            for size in self._active_sizes():
                group = self.templates_by_effective_size[size]
                r, d, m = self._get_jess_parameters(size)
                self._groups.append((size, len(ordered), len(ordered) + len(group)))
                ordered.extend(group)
                rmsd.extend([r] * len(group))
                dist.extend([d] * len(group))
                dyn.extend([m] * len(group))
            if not ordered:
                raise ValueError("no templates to search with")
            self._ordered = ordered
            self._compiled = CompiledLibrary(ordered, rmsd, dist, dyn)
        return self._compiled

    def _ensure_engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self._compile(), self.device)
        return self._engine

    def _submit(self, session, batch, stream: int = 0) -> None:
        """Upload one packed batch and enqueue every size group's search on ``stream``."""
        cutoff = float(self.conservation_cutoff) if (self.apply_conservation_mask and self.conservation_cutoff) else 0.0
        session.upload(batch, stream=stream)
        common = dict(max_candidates=self.max_candidates, ignore_chain=True, conservation_cutoff=cutoff, stream=stream)
        if self.skip_smaller_hits:
            # one launch per size group; the device skips structures that already hold a surviving hit
            mode = 1 if self.filter_matches else 2
            for gi, (_, lo, hi) in enumerate(self._groups):
                session.run(template_begin=lo, template_end=hi, skip_mode=mode, reset=(gi == 0), **common)
        else:
            session.run(template_begin=0, template_end=len(self._ordered), skip_mode=0, reset=True, **common)

    def _hit_capacity(self, n_structures: int) -> int:
        return max(self._hit_floor, self.hits_per_structure * n_structures)

    def _search(self, batch) -> np.ndarray:
        """One packed batch through the device: every size group, hit records sorted by
        (structure, template index).  A hit buffer that turns out too small (many hits per
        structure, e.g. unfiltered runs with loose cutoffs) is enlarged and the batch run again."""
        engine = self._ensure_engine()
        capacity = self._hit_capacity(batch.n_structures)
        for _ in range(6):
            session = engine.session_for(batch.n_atoms, batch.n_structures, capacity)
            self._submit(session, batch)
            try:
                return session.download()
            except EngineError as exc:
                if exc.status == -5 and exc.hits is not None:
                    # structures outside the engine's input contract were skipped on the device; the
                    # rest of the batch was searched -- keep those results instead of losing the batch
                    warnings.warn(f"{len(exc.bad_structures)} structure(s) were not searched (index: status "
                                  f"{exc.bad_structures}; 2 = more than 4194303 atoms, 3 = a residue with more "
                                  f"than 1023 atoms): {exc}")
                    self.skipped_structures = dict(exc.bad_structures)
                    return exc.hits
                if exc.status != -4:
                    raise
                capacity = 4 * session.hit_capacity
                self.hits_per_structure = max(self.hits_per_structure, capacity // max(batch.n_structures, 1))
        raise EngineError(-4, "hit buffer kept overflowing")

    def scan_files(self, paths: Sequence[os.PathLike], chunk_size: int = 2048, threads: int = 0, queue=None,
                   with_batch: bool = False, devices: Optional[Sequence[int]] = None, _with_span: bool = False,
                   on_error: str = "raise"):
        """Screen PDB / mmCIF files (gzip-compressed or not) without building ``Molecule`` objects: a generator of
        ``(chunk_paths, header_ids, records)`` per chunk of ``chunk_size`` files, ``records`` being the
        hit records of the chunk (``structure`` indexes ``chunk_paths``; ``flags & 4`` = passes the
        filter).  Three stages overlap: files are read, parsed and packed natively
        (``packing.pack_files``) on a background thread; the packed chunk crosses PCIe on one CUDA
        stream while the previous chunk is searched on another (two device sessions).
        ``matches_for`` turns the records of one file into the ``Match`` objects ``run`` would have
        returned for it.  With ``queue`` (a ``sharding.ChunkQueue`` over ``len(paths)``) the chunks
        are not taken in order but pulled from the queue, which several ranks -- one per GPU --
        share: each file is searched by exactly one of them.

        ``devices=[0, 1, ...]`` is the whole-box call from ONE process: a worker thread per listed GPU
        (its own device library, sessions and streams; the compiled library is shared) pulls chunks
        from one in-process counter and the generator hands the chunks back in input order -- the host
        side merge of north_star; nothing crosses between the GPUs.

        ``on_error="skip"``: a file that cannot be read or parsed is reported with a warning and has no
        hits instead of ending the scan (``packing.pack_files``); the default raises, as the reference's
        ``load_molecules`` would.

        Paths that all end in ``.emmpack`` are packed corpus files (``packing.write_corpus``) and go
        through ``scan_corpus`` instead: no text is parsed."""
        from .packing import is_corpus
        paths = list(paths)
        if paths and all(is_corpus(p) for p in paths):
            if devices is not None and len(list(devices)) > 0 and not (len(list(devices)) == 1 and list(devices)[0] == self.device):
                if queue is not None:
                    raise ValueError("devices= and queue= are two ways to share one list: pass one of them")
                yield from self._scan_devices(paths, list(devices), chunk_size, threads, with_batch, on_error, corpus=True)
            else:
                yield from self.scan_corpus(paths, chunk_size, with_batch=with_batch, queue=queue, _with_span=_with_span)
            return
        if devices is not None and len(list(devices)) > 0 and not (len(list(devices)) == 1 and list(devices)[0] == self.device):
            yield from self._scan_devices(paths, list(devices), chunk_size, threads, with_batch, on_error)
            return
        import concurrent.futures
        from .engine import Session
        from .packing import pack_files
        if not self._active_sizes():
            return
        engine = self._ensure_engine()
        paths = [os.fspath(p) for p in paths]
        spans = iter(queue) if queue is not None else iter([(i, min(i + chunk_size, len(paths)))
                                                             for i in range(0, len(paths), chunk_size)])
        first = next(spans, None)
        if first is None:
            return
        if self._scan_lanes is None:        # [session, stream] x 2, kept for the next call: no reallocation per scan
            self._scan_lanes = [[None, engine.new_stream()], [None, engine.new_stream()]]
        lanes = self._scan_lanes

        def collect(lane, span, chunk, ids, batch):
            extra = ((batch,) if with_batch else ()) + ((span,) if _with_span else ())
            try:
                return (chunk, ids, lane[0].download(stream=lane[1])) + extra
            except EngineError as exc:
                if exc.status == -5 and exc.hits is not None:       # refused structures: keep the rest of the chunk
                    warnings.warn(f"{len(exc.bad_structures)} structure(s) of a chunk were not searched: {exc}")
                    return (chunk, ids, exc.hits) + extra
                if exc.status != -4:
                    raise
            return (chunk, ids, self._search(batch)) + extra       # rare: rerun this chunk alone with a larger hit buffer

        in_flight: collections.deque = collections.deque()      # (lane, span, chunk paths, ids, batch)
        try:
            with concurrent.futures.ThreadPoolExecutor(max_workers=1) as pool:
                chunk = paths[first[0]:first[1]]
                pending = pool.submit(pack_files, chunk, engine.compiled, True, threads, False, on_error)
                ci = -1
                span = first
                while pending is not None:
                    ci += 1
                    batch, ids = pending.result()
                    this_chunk, this_span = chunk, span
                    span = next(spans, None)
                    if span is not None:
                        chunk = paths[span[0]:span[1]]
                        pending = pool.submit(pack_files, chunk, engine.compiled, True, threads, False, on_error)
                    else:
                        pending = None
                    lane = lanes[ci % 2]
                    sess = lane[0]
                    need_hits = self._hit_capacity(batch.n_structures)
                    if sess is None or batch.n_atoms > sess.max_atoms or batch.n_structures > sess.max_structures \
                            or need_hits > sess.hit_capacity:
                        if sess is not None:
                            sess.close()
                        grow = lambda v, old: max(int(v * 1.1) + 1, old)
                        sess = lane[0] = Session(engine.device_library,
                                                 grow(batch.n_atoms, sess.max_atoms if sess else 0),
                                                 grow(batch.n_structures, sess.max_structures if sess else 0), need_hits)
                    self._submit(sess, batch, stream=lane[1])
                    in_flight.append((lane, this_span, this_chunk, ids, batch))
                    if len(in_flight) == 2:
                        yield collect(*in_flight.popleft())
                while in_flight:
                    yield collect(*in_flight.popleft())
        finally:
            while in_flight:                    # generator abandoned early: let the device finish, drop the hits
                lane = in_flight.popleft()[0]
                try:
                    lane[0].download(stream=lane[1])
                except EngineError:
                    pass

    @staticmethod
    def corpus_plan(corpus_paths: Sequence[os.PathLike], chunk_size: int = 4096) -> List[Tuple[int, int, int, int]]:
        """The chunks of a list of corpus files, in input order: ``(file index, first structure, one past the
        last, position of the first structure in the whole list)`` -- what ``scan_corpus`` iterates over and
        what a shared ``queue`` hands out (``ChunkQueue(len(plan), 1)``)."""
        from .packing import Corpus
        plan, start = [], 0
        for k, path in enumerate(corpus_paths):
            n = Corpus.size_of(path)
            plan += [(k, lo, min(lo + chunk_size, n), start + lo) for lo in range(0, n, chunk_size)]
            start += n
        return plan

    def scan_corpus(self, corpus_paths: Sequence[os.PathLike], chunk_size: int = 4096, with_batch: bool = False,
                    queue=None, _with_span: bool = False):
        """Screen packed corpus files (``packing.write_corpus``: structures parsed and packed once, kept
        as the raw columns of the upload): a generator of ``(query_ids, header_ids, records)`` per chunk of
        ``chunk_size`` structures, as ``scan_files`` yields them for text files.  The files are mapped;
        a chunk's typing classes are expanded from the stored kinds for this matcher's library on a
        background thread while the previous chunk is on the device; nothing is parsed (SURVEY.md 8f-2:
        at > 10^4 structures/s the text parse is the wall).  ``queue``: an iterable of ``(i, j)`` index
        ranges into ``corpus_plan(corpus_paths, chunk_size)`` shared by several workers or ranks
        (``sharding.ChunkQueue(len(plan), 1)``), instead of the whole plan in order."""
        import concurrent.futures
        from .packing import Corpus
        if not self._active_sizes():
            return
        engine = self._ensure_engine()
        corpus_paths = [os.fspath(p) for p in corpus_paths]
        plan = self.corpus_plan(corpus_paths, chunk_size)
        order = (k for i, j in queue for k in range(i, j)) if queue is not None else iter(range(len(plan)))
        opened: Dict[int, Corpus] = {}                    # at most two files stay mapped: this one and the next

        def produce(index):
            k, lo, hi, start = plan[index]
            corpus = opened.get(k)
            if corpus is None:
                for stale in [f for f in opened if f < k - 1]:
                    del opened[stale]
                corpus = opened[k] = Corpus(corpus_paths[k])
            return corpus.chunk(lo, hi, engine.compiled), corpus.ids(lo, hi), (start, start + hi - lo)

        def chunks():
            with concurrent.futures.ThreadPoolExecutor(max_workers=1) as pool:
                first = next(order, None)
                pending = pool.submit(produce, first) if first is not None else None
                while pending is not None:
                    batch, ids, span = pending.result()
                    following = next(order, None)
                    pending = pool.submit(produce, following) if following is not None else None
                    yield batch, ids, span

        # two device sessions on two streams, as scan_files: chunk i+1 crosses PCIe while chunk i is searched
        from .engine import Session
        if self._scan_lanes is None:
            self._scan_lanes = [[None, engine.new_stream()], [None, engine.new_stream()]]
        lanes = self._scan_lanes

        def collect(lane, batch, ids, span):
            headers = batch.header_ids
            extra = ((batch,) if with_batch else ()) + ((span,) if _with_span else ())
            try:
                return (ids, headers, lane[0].download(stream=lane[1])) + extra
            except EngineError as exc:
                if exc.status == -5 and exc.hits is not None:       # refused structures: keep the rest of the chunk
                    warnings.warn(f"{len(exc.bad_structures)} structure(s) of a chunk were not searched: {exc}")
                    return (ids, headers, exc.hits) + extra
                if exc.status != -4:
                    raise
            return (ids, headers, self._search(batch)) + extra     # rare: rerun this chunk alone with a larger hit buffer

        in_flight: collections.deque = collections.deque()          # (lane, batch, ids, span)
        try:
            for ci, (batch, ids, span) in enumerate(chunks()):
                lane = lanes[ci % 2]
                sess = lane[0]
                need_hits = self._hit_capacity(batch.n_structures)
                if sess is None or batch.n_atoms > sess.max_atoms or batch.n_structures > sess.max_structures \
                        or need_hits > sess.hit_capacity:
                    if sess is not None:
                        sess.close()
                    grow = lambda v, old: max(int(v * 1.1) + 1, old)
                    sess = lane[0] = Session(engine.device_library,
                                             grow(batch.n_atoms, sess.max_atoms if sess else 0),
                                             grow(batch.n_structures, sess.max_structures if sess else 0), need_hits)
                self._submit(sess, batch, stream=lane[1])
                in_flight.append((lane, batch, ids, span))
                if len(in_flight) == 2:
                    yield collect(*in_flight.popleft())
            while in_flight:
                yield collect(*in_flight.popleft())
        finally:
            while in_flight:                    # generator abandoned early: let the device finish, drop the hits
                lane = in_flight.popleft()[0]
                try:
                    lane[0].download(stream=lane[1])
                except EngineError:
                    pass

    def _scan_devices(self, paths, devices: List[int], chunk_size: int, threads: int, with_batch: bool,
                      on_error: str = "raise", corpus: bool = False):
        """``scan_files`` over several GPUs of this host from one process (see ``scan_files``).  ``corpus``:
        the paths are packed corpus files and the workers share the chunk plan instead of the path list --
        no text is parsed, so one process can keep every GPU of the box busy."""
        import queue as queue_module
        import threading
        paths = [os.fspath(p) for p in paths]
        self._compile()
        if corpus:
            spans = iter([(i, i + 1) for i in range(len(self.corpus_plan(paths, chunk_size)))])
        else:
            spans = iter([(i, min(i + chunk_size, len(paths))) for i in range(0, len(paths), chunk_size)])
        lock = threading.Lock()

        class Shared:                           # one counter for all workers: dynamic hand-out of chunks
            def __iter__(self):
                return self

            def __next__(self):
                with lock:
                    return next(spans)

        if not hasattr(self, "_device_workers"):
            self._device_workers = {}
        host = _available_cpus()
        per_worker = threads or max(1, host // len(devices))
        results: "queue_module.Queue" = queue_module.Queue(maxsize=4 * len(devices))
        stop = threading.Event()                # set when the consumer goes away before the scan is over

        def hand_over(item) -> bool:
            while not stop.is_set():
                try:
                    results.put(item, timeout=0.2)
                    return True
                except queue_module.Full:
                    continue
            return False

        def work(slot, device):
            try:
                child = self._device_workers.get((slot, device))
                if child is None:
                    child = Matcher.__new__(Matcher)
                    child.__dict__.update(self.__dict__)
                    child.device, child._engine, child._scan_lanes, child._device_workers = device, None, None, {}
                    self._device_workers[(slot, device)] = child
                scan = child.scan_corpus(paths, chunk_size, with_batch=with_batch, queue=Shared(), _with_span=True) if corpus \
                    else child.scan_files(paths, chunk_size, per_worker, queue=Shared(), with_batch=with_batch,
                                          _with_span=True, on_error=on_error)
                for item in scan:
                    if not hand_over(item):
                        break                   # closing the inner generator lets its device work finish
            except BaseException as exc:        # noqa: BLE001 -- surfaces in the consumer
                hand_over(exc)
            finally:
                hand_over(None)

        workers = [threading.Thread(target=work, args=(i, d), daemon=True) for i, d in enumerate(devices)]
        for w in workers:
            w.start()
        waiting, finished, next_start = {}, 0, 0
        try:
            while finished < len(workers):
                item = results.get()
                if item is None:
                    finished += 1
                    continue
                if isinstance(item, BaseException):
                    raise item
                waiting[item[-1][0]] = item[:-1]
                while next_start in waiting:        # hand chunks back in input order
                    out = waiting.pop(next_start)
                    next_start += len(out[0])
                    yield out
        finally:
            stop.set()                              # abandoned or failed: workers stop after their current chunk
            for w in workers:
                w.join()

    def scan_to_tsv(self, paths: Sequence[os.PathLike], file: IO[str], chunk_size: int = 2048, threads: int = 0,
                    queue=None, header: bool = True, predict_correctness: bool = True,
                    devices: Optional[Sequence[int]] = None, on_error: str = "raise") -> int:
        """PDB files -> the reference's results table, end to end: native ingest, GPU search, rows
        formatted natively from the hit records (``tsv.TableWriter``) -- what ``_cli.py:217-316`` does
        through ``load_molecules`` / ``Matcher.run`` / ``Match.dump``, without per-atom or per-match
        Python objects.  Query ids are the file stems (repeats get ``_2``, ``_3`` ... as in
        ``load_molecules``).  Rows are ordered as the reference orders them within each chunk of
        ``chunk_size`` files, chunks in input order -- the order of the reference's own scale-out,
        which splits the list and concatenates the tables (``nextflow/template_matcher.nf:43-62``).
        Returns the number of rows written."""
        from .tsv import TableWriter
        writer = TableWriter(self, predict_correctness)
        paths = [os.fspath(p) for p in paths]
        seen: Dict[str, int] = collections.defaultdict(int)
        stems = {}
        for path in paths:
            stem = Path(path).stem
            seen[stem] += 1
            stems[path] = stem if seen[stem] == 1 else f"{stem}_{seen[stem]}"
        # the rows arrive as UTF-8 bytes: written as they are to a binary file, decoded for a text file
        binary = isinstance(file, (io.RawIOBase, io.BufferedIOBase)) or "b" in str(getattr(file, "mode", ""))
        emit = (lambda b: file.write(b)) if binary else (lambda b: file.write(b.decode("utf-8")))
        if header:
            emit(writer.header().encode())
        n_rows = 0
        from .packing import is_corpus
        stored_ids = bool(paths) and all(is_corpus(p) for p in paths)       # corpus files carry their query ids
        for chunk, _, records, batch in self.scan_files(paths, chunk_size, threads, queue, with_batch=True, devices=devices,
                                                        on_error=on_error):
            selection = writer.select(records)
            n_rows += len(selection[0])
            emit(writer.format(records, batch.table, list(chunk) if stored_ids else [stems[p] for p in chunk], selection))
        return n_rows

    def run_to_tsv(self, molecules: List[Molecule], file: IO[str], header: bool = True,
                   predict_correctness: bool = True) -> int:
        """``Matcher.run`` + ``Match.dump`` of every match (``_cli.py:248-316``) in one call, rows
        formatted natively.  Returns the number of rows written."""
        from .tsv import TableWriter
        writer = TableWriter(self, predict_correctness)
        if header:
            file.write(writer.header())
        if not self._active_sizes() or not molecules:
            return 0
        engine = self._ensure_engine()
        batch = pack_molecules(molecules, engine.compiled)
        if batch.table is None:
            raise ValueError("run_to_tsv needs molecules read by Molecule.load / load_molecules")
        records = self._search(batch)
        selection = writer.select(records)
        file.write(writer.format(records, batch.table, [m.id for m in molecules], selection).decode("utf-8"))
        return len(selection[0])

    def close(self) -> None:
        """Release the device sessions, streams and library of this matcher."""
        engine = self._engine
        if self._scan_lanes is not None and engine is not None:
            for sess, stream in self._scan_lanes:       # destroying a session waits for its device work
                if sess is not None:
                    sess.close()
                engine.free_stream(stream)
        self._scan_lanes = None
        for child in getattr(self, "_device_workers", {}).values():
            child.close()
        self._device_workers = {}
        if engine is not None:
            engine.close()
            self._engine = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def matches_for(self, molecule: Molecule, records: np.ndarray) -> List[Match]:
        """``Match`` objects for the hit records of one structure (``records`` = the rows of a
        ``scan_files`` chunk whose ``structure`` is this molecule), as ``run`` builds them."""
        relabel = records.copy()
        relabel["structure"] = 0
        return self._assemble(relabel, [molecule]).get(molecule, [])

    def run(self, molecules: List[Molecule]) -> Dict[Molecule, List[Match]]:
        """Search every molecule against every size group; ``{molecule: [Match, ...]}``."""
        processed: Dict[Molecule, List[Match]] = collections.defaultdict(list)
        if not self._active_sizes() or not molecules:
            return processed
        engine = self._ensure_engine()
        return self._assemble(self._search(pack_molecules(molecules, engine.compiled)), molecules)

    def _template_identity(self):
        """Per compiled template: (completeness group code or -1, cluster member, cluster size) --
        ``Match.get_identifying_attributes`` / ``_check_completeness`` (``jess_run.py:286-296, 738-783``)
        evaluated once per template instead of once per hit."""
        cached = getattr(self, "_identity", None)
        if cached is None or len(cached[0]) != len(self._ordered):
            codes: Dict[tuple, int] = {}
            ident, member, size = [], [], []
            for t in self._ordered:
                if t.mcsa_id is not None and t.cluster is not None:
                    ident.append(codes.setdefault((t.mcsa_id, t.cluster.id, t.dimension), len(codes)))
                    member.append(t.cluster.member)
                    size.append(t.cluster.size)
                else:
                    ident.append(-1)
                    member.append(0)
                    size.append(0)
            cached = self._identity = (ident, member, size)
        return cached

    def _assemble(self, records: np.ndarray, molecules: Sequence[Molecule]) -> Dict[Molecule, List[Match]]:
        """Hit records -> ``{molecule: [Match]}`` with the reference's per-size-group completeness
        check and filtering (``jess_run.py:845-894``).  Which records survive, their completeness and
        their order are array operations (``tsv.RowSelector``, shared with the table writer); Python
        objects are made for the surviving matches only, and a ``Hit`` decodes its atoms and
        transform when asked."""
        from .tsv import RowSelector
        processed: Dict[Molecule, List[Match]] = collections.defaultdict(list)
        selector = getattr(self, "_selector", None)
        if selector is None or selector.bounds is None or len(selector.ident) != len(self._ordered):
            selector = self._selector = RowSelector(self, predict_correctness=False)
        try:
            rows, _, complete, _ = selector.select(records)
        except KeyError:
            # a hit at a distance without logistic models: raise what the reference raises, from the
            # reference's own formula (jess_run.py:339-342)
            bad = records[(records["flags"] & 0x08) != 0][0]
            template = self._ordered[int(bad["template_index"])]
            distance = self._get_jess_parameters(template.effective_size)[1]
            Match(hit=Hit(bad, template, molecules[int(bad["structure"])]), pairwise_distance=distance).predicted_correct
            raise
        rows_l = rows.tolist()
        picked = records[rows] if len(rows) else records[:0]
        cols = zip(picked["rmsd"].tolist(), picked["orientation"].tolist(), picked["flags"].tolist(),
                   picked["n_complete"].tolist(), picked["template_index"].tolist(), picked["structure"].tolist())
        ordered = self._ordered
        distance_of = np.empty(len(ordered), dtype=object)
        for gsize, lo, hi in self._groups:
            distance_of[lo:hi] = self._get_jess_parameters(gsize)[1]
        overflowed = 0
        for i, done, sc in zip(rows_l, complete.tolist(), cols):
            molecule = molecules[sc[5]]
            processed[molecule].append(Match(hit=Hit(records[i], ordered[sc[4]], molecule, sc), complete=bool(done),
                                             pairwise_distance=distance_of[sc[4]]))
            overflowed += sc[2] & 0x01
        if self.verbose:
            group_of = np.searchsorted(selector.bounds, picked["template_index"], side="right") if len(rows) else []
            seen = 0
            for gi, (gsize, lo, hi) in enumerate(self._groups):
                rmsd, distance, max_dyn = self._get_jess_parameters(gsize)
                self.verbose_print(f"Now matching query structure(s) to template of size {gsize}")
                self.verbose_print(f"jess parameters are: {rmsd} {distance} {max_dyn}")
                in_group = picked["structure"][np.asarray(group_of) == gi] if len(rows) else []
                self.verbose_print(f"{len(in_group)} matches found!")
                seen = len(set(picked["structure"][np.asarray(group_of) <= gi].tolist())) if len(rows) else 0
                self.verbose_print(f"{seen} target structures processed!")
        if overflowed and self.warn:
            # at the cap the best hit comes from the candidates examined so far, in this
            # engine's enumeration order -- not reproducible against Jess; Match.hit.overflow marks them
            warnings.warn(f"{overflowed} match(es) reached max_candidates={self.max_candidates}: their template has more "
                          "candidate assignments than were examined, so a better one may exist (Match.hit.overflow)")
        return processed

    def run_single(self, molecule: Molecule) -> List[Match]:
        return self.run([molecule])[molecule]

    # signature-compatible single (molecule, size group) entry point (``jess_run.py:785-843``)
    @staticmethod
    def _run_jess(molecule: Molecule, templates: List[Template], rmsd_threshold: float = 2.0,
                  distance_cutoff: float = 1.5, max_dynamic_distance: float = 1.5,
                  max_candidates: int = 10000) -> List[Match]:
        query = pyjess.Jess(templates).query(molecule, rmsd_threshold, distance_cutoff, max_dynamic_distance,
                                             max_candidates=max_candidates, best_match=True, ignore_chain=True)
        return [Match(hit=hit, pairwise_distance=distance_cutoff) for hit in query]

    def _single_query_run(self, molecule, templates, rmsd_threshold, distance_cutoff, max_dynamic_distance,
                          max_candidates: int = 10000) -> List[Match]:
        return self._check_completeness(self._run_jess(molecule, templates, rmsd_threshold, distance_cutoff,
                                                       max_dynamic_distance, max_candidates))
