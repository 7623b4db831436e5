import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json

    with open(os.path.join(ROOT, "tests", "golden", "classic_prions.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session", autouse=True)
def _built_artifacts():
    """The CUDA library, host CLI and oracle are built in-tree (git-ignored); build them if this checkout is fresh.
    nvcc cross-compiles without a GPU."""
    from oracle import orc
    from plaac_b200 import build as pb

    pb.build_lib()
    pb.build_cli()
    orc.build()
