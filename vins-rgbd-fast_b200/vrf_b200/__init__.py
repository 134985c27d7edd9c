"""Harness-side python package for the B200-native VINS-RGBD-FAST hot path.

The product is the C-ABI shared library (include/vrf.h, csrc/); this package
only holds the ctypes binding used by tests/bench and the synthetic-sequence
generator.  Nothing here computes on the hot path.
"""
