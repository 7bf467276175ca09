"""kmerpapa_b200 — B200-native optimal k-mer pattern partition (the hot path of kmerPaPa).

The dynamic program runs in hand-written sm_100a CUDA kernels behind a C ABI (libkpapa.so, declared
in include/kmerpapa_b200.h); this package is the host side: it mirrors the two reference functions
cli.py calls (kmerpapa_b200.algorithms.*) and keeps the `kmerpapa` command line unchanged.
There is no CPU implementation of the DP in this package: without the built library and a CUDA
device every entry point raises.
"""
__version__ = "0.2.4+b200.1"
