import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json

    import numpy as np
    data = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    index = json.loads(bytes(data["index_json"]).decode())
    return [(meta, data["in_%d" % i], data["out_%d" % i]) for i, meta in enumerate(index)]


@pytest.fixture(scope="session")
def golden_ops():
    """Reference outputs of the compressed-domain operations (tools/gen_golden_ops.py)."""
    import json

    import numpy as np
    data = np.load(os.path.join(ROOT, "tests", "golden", "golden_ops_v1.npz"))
    index = json.loads(bytes(data["index_json"]).decode())
    return [(meta, data["in_%d" % i], data["out_%d" % i]) for i, meta in enumerate(index)]


@pytest.fixture(scope="session")
def icb():
    """The product library.  Building it is __graft_entry__.build()'s job; here it must simply be present."""
    import image_compression_b200
    image_compression_b200.lib()
    return image_compression_b200
