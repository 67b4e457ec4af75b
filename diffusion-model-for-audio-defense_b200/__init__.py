"""audiopure_b200: B200-native (sm_100a) AudioPure purification-and-classify hot path.

Import as ``audiopure_b200`` (see ../audiopure_b200.py).  Sub-modules are imported lazily so that
``audiopure_b200.synthetic`` and ``audiopure_b200.build`` work before the CUDA library has been built; every compute
module loads ``libaudiopure_b200.so`` and fails loudly if it is missing (there is no CPU / PyTorch fallback).
"""
import importlib

__version__ = "0.1.0"

_LAZY = {
    "calc_diffusion_hyperparams": "diffwave", "WaveNet": "diffwave", "DiffWave": "diffwave", "ReffWave": "diffwave",
    "create_diffwave_model": "diffwave",
    "RevDiffWave": "diffwave_sde", "RevVPSDE": "diffwave_sde",
    "MelSpectrogramDB": "transforms", "sc09_transform": "transforms", "kws_transform": "transforms",
    "ResNeXtClassifier": "classifiers", "ResNetClassifier": "classifiers", "VGGClassifier": "classifiers", "WideResNetClassifier": "classifiers", "DenseNetClassifier": "classifiers", "M5Classifier": "classifiers", "KWSClassifier": "classifiers",
    "create_model": "classifiers",
    "AcousticSystem": "acoustic_system",
    "UNet": "improved_diffusion", "RevImprovedDiffusion": "improved_diffusion",
    "RobustCertificate": "certify", "certify_dataset": "certify",
    "EOT": "blackbox", "NES": "blackbox", "QueryLoss": "blackbox", "resolve_loss": "blackbox", "resolve_prediction": "blackbox",
    "AudioPureError": "_lib",
}


def __getattr__(name):
    if name in _LAZY:
        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
