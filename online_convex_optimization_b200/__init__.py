"""online_convex_optimization_b200 -- B200-native engine for the per-timestep hot path of
revvu/online_convex_optimization (FTL / FTRL / SMART trajectory loops and their seeded
sequence generators), behind the reference's own module interface.

    from online_convex_optimization_b200 import fast_algorithms   # drop-in for the reference module
    from online_convex_optimization_b200 import engine            # batched entry points

All compute runs in hand-written sm_100a CUDA kernels (csrc/, C ABI in include/oco_b200.h).
There is no CPU fallback: importing works anywhere, calling a compute function without a
CUDA device raises RuntimeError.
"""
__version__ = "0.1.0"
