import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def all_templates():
    from enzymm_b200.templates import load_templates
    return list(load_templates())


@pytest.fixture(scope="session")
def active_templates(all_templates):
    return [t for t in all_templates if t.effective_size >= 3]


@pytest.fixture(scope="session")
def mol_1amy():
    from enzymm_b200.structures import Molecule
    return Molecule.load(GOLDEN / "1AMY.pdb")


@pytest.fixture(scope="session")
def mol_af():
    from enzymm_b200.structures import Molecule
    return Molecule.load(GOLDEN / "AF-P0DUB6-F1-model_v4.pdb")


def find_template(templates, id_string, effective_size, cluster=None):
    for t in templates:
        if t.template_id_string == id_string and t.effective_size == effective_size:
            if cluster is None or (t.cluster.id, t.cluster.member, t.cluster.size) == cluster:
                return t
    raise KeyError((id_string, effective_size, cluster))


@pytest.fixture(scope="session")
def template_1uh3(all_templates):
    # 5_residues/results/csa3d_0285/csa3d_0285.cluster_1_1_1.1uh3_A396-A262-A356-A471-A472
    return find_template(all_templates, "1uh3_A396-A262-A356-A471-A472", 5, (1, 1, 1))


@pytest.fixture(scope="session")
def template_1be0(all_templates):
    # 3_residues/results/csa3d_0415/csa3d_0415.cluster_1_1_2.1be0_A124-A175-A125-A289-A260
    return find_template(all_templates, "1be0_A124-A175-A125-A289-A260", 3, (1, 1, 2))
