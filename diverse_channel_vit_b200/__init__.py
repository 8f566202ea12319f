"""B200-native (sm_100a) implementation of the DiChaViT training hot path.

Drop-in for the reference's models/dichavit.py `DiChaViT` module; all compute runs in
hand-written CUDA kernels behind the C ABI of libdcvit.so (include/dcvit.h).
"""
__version__ = "0.1.0"
