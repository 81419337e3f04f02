"""pesr_b200: B200-native (sm_100a) implementation of the PESR training / inference hot path.

Drop-in surface: ``pesr_b200.model`` mirrors the reference's ``model`` package (Generator,
Discriminator, VGG, FocalLoss); everything under it runs on hand-written CUDA kernels behind the
C-ABI declared in ``include/pesr_b200.h``.
"""
__version__ = "0.1.0"
