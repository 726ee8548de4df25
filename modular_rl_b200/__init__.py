"""modular_rl_b200 - B200-native policy-update path of modular_rl (TRPO / PPO-LBFGS + GAE).

Python host code mirrors the reference's operator interface (TrpoUpdater, PpoLbfgsUpdater,
compute_advantage, NnVf, ZFilter, agentzoo agents) and drives hand-written sm_100a kernels
through the C ABI in include/mrl_b200.h.  Importing this package does not load the CUDA
library; the first operator call does, and raises if it is not built (no CPU fallback).
"""
__version__ = "0.1.0"
