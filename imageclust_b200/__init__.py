"""imageclust_b200 -- B200-native size-constrained Ward clustering.

Drop-in for ``internal/clustering.PerformClusteringWithConstraints`` of
monahand1023/imageclust (``/root/reference/internal/clustering/clustering.go:198``),
implemented as hand-written sm_100a CUDA kernels behind a C ABI
(``include/imageclust_b200.h``).  There is no CPU fallback: importing
``imageclust_b200.clustering`` works anywhere, calling it without the built
CUDA library or without a GPU raises.
"""
__version__ = "0.1.0"
