#!/usr/bin/env python
"""Device time of the sketch kernels on long ragged reads (BASELINE config 5 in miniature: 10 kbp reads with N runs,
k=31, s=11), with and without the on-device re-tiling (NTC_NO_RETILE=1 forces the general kernel).
    python tools/bench_longreads.py [n_reads]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntcard_b200 as nt  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    L, k, sBits = 10000, 31, 11
    chars = nt.gen_ascii(4, 0, n, L, mode=2)
    soff = np.arange(n + 1, dtype=np.uint64) * L
    words, off = nt.pack_chars(chars, soff, min_len=k)
    kmers = None
    for tag in ("retiled", "general"):
        with nt.Sketch([k], rBits=27, sBits=sBits) as sk:
            if tag == "general":
                sk.set_kernel(nt.KERNEL_ROLL64)
            for _ in range(2):
                sk.reset()
                sk.submit(words, off)
                sk.sync()
            sk.kernel_time()
            reps = 3
            for _ in range(reps):
                sk.reset()
                sk.submit(words, off)
                sk.flush()
            sk.sync()
            ms, _ = sk.kernel_time()
            kmers = int(sk.totals()[0]) // 1   # of the last repetition
            print(f"{tag}: {ms / reps:.3f} ms per pass, {kmers / (ms / reps * 1e-3):.3e} k-mers/s ({n} reads x {L} bp, {len(off) - 1} records, k={k}, s={sBits})",
                  flush=True)


if __name__ == "__main__":
    main()
