"""egopose_b200: B200-native PPO rollout-and-update hot path behind EgoPose's Agent/Policy/Value API.

Everything numerical runs in hand-written sm_100a CUDA kernels exported by a C-ABI shared library
(include/egopose_b200.h, built by egopose_b200.build); there is no CPU fallback.
"""
__version__ = '0.1.0'
