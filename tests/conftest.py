import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "graphs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def fixture_graphs():
    """All fixture graph names under tests/golden/graphs (from X.properties)."""
    return sorted(f[: -len(".properties")] for f in os.listdir(GOLDEN) if f.endswith(".properties"))


def golden_cases():
    """(graph, ALG) for each of the 24 golden output files."""
    cases = []
    for g in fixture_graphs():
        for alg in ("BFS", "CDLP", "LCC", "PR", "SSSP", "WCC"):
            if os.path.exists(os.path.join(GOLDEN, f"{g}-{alg}")):
                cases.append((g, alg))
    return cases


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
