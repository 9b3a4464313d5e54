import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def model_path(name):
    return os.path.join(GOLDEN, "models", name + ".xml")


def is_membrane_species(g, sp):
    return ("memgeo/%s" % sp) in g.files


def species_values(sim, g, sp):
    """The values of a species laid out like the golden file stores them: compact volume cells, or - for a species that lives on
    a membrane - its values at every classified point of that membrane (geo/<membrane>/points, doubled-grid indices)."""
    if is_membrane_species(g, sp):
        pts = g["geo/%s/points" % str(g["memgeo/%s" % sp])]
        return np.asarray(sim.getVariable(sp))[pts]
    return sim.getVariableCompact(sp)


@pytest.fixture(scope="session")
def lib():
    import spatial_sbml_b200 as sb
    return sb.load_library()
