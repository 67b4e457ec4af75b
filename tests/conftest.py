import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def golden_grad():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_golden_grad.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def golden_blackbox():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_golden_blackbox.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def golden_vgg():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_golden_vgg.npz")
    return dict(np.load(path))
