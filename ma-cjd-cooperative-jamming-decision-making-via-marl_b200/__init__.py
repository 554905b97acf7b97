"""B200-native hot path of MA-CJD: batched environment step + QMix/MP-DQN
act-and-learn loop, behind the reference's own Python class API.

Sub-modules mirror the reference layout (simulation/, core/, utils/, runners/);
``csrc/`` holds the sm_100a CUDA kernels and the C-ABI library (include/macjd.h).
"""
__version__ = "0.1.0"
