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


@pytest.fixture(scope="session")
def golden_v2():
    """outputs of the unmodified reference at the BASELINE shapes (tests/golden/make_golden_v2.py)"""
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_v2.npz")))


@pytest.fixture(scope="session")
def golden_sde():
    """the reference's own RevDiffWave / RevVPSDE driven through a restated Euler loop (tests/golden/make_golden_sde.py)"""
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_sde.npz")))


@pytest.fixture(scope="session")
def trained_checkpoints():
    """state dicts of the trained M5 / RCNN_KWS checkpoints the reference ships: {'m5': {...}, 'kws': {...}}"""
    import numpy as np
    raw = np.load(os.path.join(ROOT, "tests", "golden", "reference_checkpoints.npz"))
    out = {"m5": {}, "kws": {}}
    for k in raw.files:
        fam, name = k.split(".", 1)
        out[fam][name] = raw[k]
    return out


@pytest.fixture(scope="session")
def golden_unet():
    """the reference's spectrogram UNet and RevImprovedDiffusion (tests/golden/make_golden_unet.py)"""
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_unet.npz")))
