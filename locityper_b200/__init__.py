"""B200-native genotype-evaluation hot path of locityper (prefilter + read-assignment solvers).

The compute path lives in `csrc/` (hand-written sm_100a CUDA behind the C ABI of `include/lctp.h`);
this package is the Python host-side mirror of the reference's interface for that path.
"""
__version__ = "0.1.0"
