"""cupss_b200 -- B200-native (sm_100a) engine for cuPSS's per-timestep integration loop.

Layout: ``csrc/`` hand-written CUDA kernels + the C ABI (``include/cupss_b200.h``), ``host/`` the C++ mirror of the
reference's evolver/parser/field/term API, ``capi.py`` a ctypes binding used by tests and bench.py.
"""
