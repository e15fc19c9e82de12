"""rnacode_b200 -- B200 (sm_100a) implementation of RNAcode's scoring hot path.

The product is the C-ABI shared library ``lib/libRNAcode_cuda.so`` (sources in ``csrc/``, interface in
``include/rnacode_cuda.h``).  This package only holds the build recipe, a thin ctypes binding used by
the tests and ``bench.py``, and the synthetic-alignment generator of SURVEY.md section 8(d).
There is no CPU fallback: importing :mod:`rnacode_b200.capi` without the built library raises.
"""

__all__ = ["build", "capi", "synth"]
