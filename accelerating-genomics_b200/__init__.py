"""accelerating-genomics_b200 -- B200 (sm_100a) Smith-Waterman / PairHMM hot paths.

The product is libagx.so (CUDA kernels + C ABI, include/agx.h) and the C drivers under drivers/.
This Python package is the thin host-side mirror used by tests and bench.py:

    capi     ctypes binding of the C ABI
    formats  the reference's text formats -> flat (buffer, offsets, lengths) arrays
    synth    seeded synthetic inputs of the BASELINE.json shapes
    build    the nvcc / gcc build recipe

The directory name carries a hyphen, so import it through the repo-root helper:
    import agxpkg; agx = agxpkg.load()
"""
from . import capi, formats, synth  # noqa: F401

__all__ = ["capi", "formats", "synth"]
