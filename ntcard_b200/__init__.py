"""ntcard_b200 -- the ntCard streaming k-mer sketch path, rebuilt for NVIDIA B200 (sm_100a).

The product is the C-ABI shared library (include/ntcard_b200.h, ntcard_b200/libntcard_b200.so) and
the `ntcard` command line built on it; this package is the thin ctypes mirror used by tests,
bench.py and Python callers.  Importing it requires the built CUDA library -- there is no CPU path.
"""
from .api import (KERNEL_AUTO, KERNEL_BITSLICE, KERNEL_ROLL64, LIB_PATH, SYMBOLS, HllSketch, NtcError, PinnedBuffer, Sketch,  # noqa: F401
                  apply_sbits_rule, check_offsets, device_count, estimate, gen_ascii, gen_packed, hll_estimate, lib, pack_chars, pack_reads,
                  stride_words, write_hist)

__all__ = ["Sketch", "HllSketch", "hll_estimate", "PinnedBuffer", "NtcError", "estimate", "pack_reads", "pack_chars", "gen_ascii", "gen_packed",
           "stride_words", "apply_sbits_rule", "check_offsets", "device_count", "write_hist", "lib", "SYMBOLS", "LIB_PATH",
           "KERNEL_AUTO", "KERNEL_ROLL64", "KERNEL_BITSLICE"]
