"""utree_b200 -- B200-native SEARCH_GG (UTree read classifier) hot path.

The product is the C-ABI library ``utree_b200/csrc/libutree_b200.so`` (CUDA
kernels for sm_100a + C host pipeline) and the ``bin/utree-search_gg`` CLI.
This package is the thin ctypes mirror used by tests and bench.
"""
