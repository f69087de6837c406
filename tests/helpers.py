"""Test-side helpers: an oracle-backed restatement of Matcher.run's bookkeeping
(reference enzymm/jess_run.py:896-988) so GPU results can be compared at the Matcher level."""
import collections
from typing import Dict, List, Sequence

import numpy as np

import oracle

DEFAULT_PARAMS = {3: (2, 0.9, 0.9), 4: (2, 1.7, 1.7), 5: (2, 2.0, 2.0), 6: (2, 2.0, 2.0),
                  7: (2, 2.0, 2.0), 8: (2, 2.0, 2.0)}


def params_for(size: int, table=None):
    table = table or DEFAULT_PARAMS
    return table[min(max(size, 3), 8)]


def default_distances(templates) -> List[float]:
    return [params_for(t.effective_size)[1] for t in templates]


class OracleMatch:
    def __init__(self, hit: oracle.OracleHit, template, molecule, distance):
        self.hit, self.template, self.molecule, self.distance = hit, template, molecule, distance
        self.transformed = hit.transform(molecule.xyz[hit.atoms])
        self.orientation = oracle.orientation(template, self.transformed)
        self.complete = False

    @property
    def predicted_correct(self):
        return oracle.predicted_correct(self.template.effective_size, self.distance, self.hit.rmsd, self.orientation)


def check_completeness(matches: Sequence[OracleMatch]):
    groups: Dict[tuple, List[OracleMatch]] = collections.defaultdict(list)
    for m in matches:
        t = m.template
        if t.mcsa_id is not None and t.cluster is not None:
            groups[(t.mcsa_id, t.cluster.id, t.dimension)].append(m)
        else:
            m.complete = True
    for members in groups.values():
        if sorted(m.template.cluster.member for m in members) == list(range(1, members[0].template.cluster.size + 1)):
            for m in members:
                m.complete = True


def oracle_matcher_run(templates, molecules, jess_params=None, filter_matches=True, skip_smaller_hits=False,
                       match_small_templates=False, max_candidates=10000, threads=4):
    """{molecule index: [OracleMatch]} with the reference's size-major ordering."""
    by_size: Dict[int, list] = collections.defaultdict(list)
    for t in templates:
        by_size[t.effective_size].append(t)
    processed: Dict[int, List[OracleMatch]] = collections.OrderedDict()
    for size in sorted(by_size, reverse=True):
        if size < 3 and not match_small_templates:
            continue
        group = by_size[size]
        rmsd, dist, dyn = params_for(size, jess_params)
        todo = [i for i in range(len(molecules)) if not (skip_smaller_hits and i in processed)]
        if not todo:
            continue
        ot = oracle.OracleTemplates(group)
        hits = oracle.query([molecules[i] for i in todo], ot, rmsd, dist, dyn, max_candidates=max_candidates,
                            ignore_chain=True, threads=threads)
        for i, mol_hits in zip(todo, hits):
            matches = [OracleMatch(h, group[h.template_index], molecules[i], dist) for h in mol_hits]
            check_completeness(matches)
            keep = [m for m in matches if m.predicted_correct] if filter_matches else matches
            if keep:
                processed.setdefault(i, []).extend(keep)
    return processed


def svd_kabsch(t: np.ndarray, q: np.ndarray):
    """Textbook SVD Kabsch: rotate q onto t about centroids; returns (rmsd, R)."""
    tc, qc = t.mean(axis=0), q.mean(axis=0)
    H = (q - qc).T @ (t - tc)
    U, _, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(Vt.T @ U.T))
    R = Vt.T @ np.diag([1.0, 1.0, d]) @ U.T
    diff = (q - qc) @ R.T - (t - tc)
    return float(np.sqrt((diff ** 2).sum() / len(t))), R
