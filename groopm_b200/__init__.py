"""groopm_b200 -- B200-native contig-profile distance and neighbour-search engine.

A drop-in for the scipy cdist / pdist / KDTree + argmin / sort calls on GroopM's
clustering hot path (cluster.py, refine.py, som.py, torusMesh.py).  The compute
lives in ``libgmdist.so`` (hand-written sm_100a CUDA behind a C ABI,
``include/gmdist.h``); this package is the Python host side.
"""
__version__ = "0.1.0"
