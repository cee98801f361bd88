"""genometools_smax_b200 -- B200-native supermaximal-repeat scan over a
GenomeTools enhanced suffix array (drop-in for the `smax` path).

The product is the C-ABI library built from ``csrc/`` (``lib/libsmax.so``,
declared in ``include/smax.h``) and the ``smax`` tool; this package holds the
ctypes mirror of that boundary (``capi``), the in-tree build (``_build``) and
the one-process-per-GPU driver (``shard``).
"""
__version__ = "0.1.0"
