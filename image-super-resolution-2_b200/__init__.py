"""B200-native FreqFusion x4 inference (drop-in for the reference's models/team29_FreqFusion path).

Python host code + hand-written sm_100a kernels behind the C ABI declared in include/ffb200.h.
There is no CPU fallback: every forward goes through csrc/libffb200.so on a CUDA device.
"""
__version__ = "0.1.0"
