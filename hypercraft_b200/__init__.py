"""hypercraft_b200 -- B200-native chunk mesher for HyperCraft (one hot path: the client mesh task).

The product is ``csrc/libhcm.so`` (hand-written sm_100a CUDA behind the C ABI of include/hcm.h); this package is
the thin Python plumbing used by tests and bench.py: ctypes bindings (``mesher``), the registry blob
(``registry``) and input definitions (``worldgen``)."""
