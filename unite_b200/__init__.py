"""unite_b200 — B200-native (sm_100a) implementation of the UNITE training-step hot path.

Host side mirrors the reference's Python API (reddyav1/unite: src/models, src/engines); all compute is
hand-written CUDA behind the C ABI declared in include/unite_b200.h (unite_b200/lib/libunite_b200.so).
There is no CPU fallback: importing the compute modules without the built library raises.
"""
__version__ = "0.1.0"
