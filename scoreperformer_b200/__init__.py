"""scoreperformer_b200: B200-native (sm_100a) implementation of the ScorePerformer training / rendering step.

Drop-in for `scoreperformer.models` (constructors, forward signatures, output dataclasses, state_dict layout); the
arithmetic runs in hand-written CUDA kernels behind a C ABI (include/spb200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
