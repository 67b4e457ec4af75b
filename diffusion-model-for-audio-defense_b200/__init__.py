"""audiopure_b200: B200-native (sm_100a) AudioPure purification-and-classify hot path.

Import as ``audiopure_b200`` (see ../audiopure_b200.py).  Sub-modules are imported lazily so that
``audiopure_b200.synthetic`` works without the CUDA library being built.
"""
__version__ = "0.1.0"
