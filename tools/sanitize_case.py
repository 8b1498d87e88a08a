#!/usr/bin/env python
"""A small pass through every kernel of the pipeline (scan, hit staged + plain, apply zero / prefetch mode, fallback,
re-tiling, general kernel, histograms) for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
    compute-sanitizer --tool synccheck python tools/sanitize_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntcard_b200 as nt  # noqa: E402


def main():
    n, L = 3000, 150
    stride = nt.stride_words(L)
    words = nt.gen_packed(5, 0, n, L, 1, n // 4, stride)
    with nt.Sketch([32, 64], rBits=22, sBits=7) as sk:
        sk.submit(words, None, n, stride)                       # scan + staged hit
        sk.flush()                                              # apply, zero mode
        sk.submit(words, None, n, stride)
        long_words = nt.gen_packed(6, 0, 1100, 400, 0, 0, nt.stride_words(400))
        sk.submit(long_words, None, 1100, nt.stride_words(400))  # plain hit kernel
        mixed = words.copy()
        mixed[0::stride][:1500:7] = 100                          # some records shorter: flagged tiles -> fallback kernel
        sk.submit(mixed, None, n, stride)
        chars = nt.gen_ascii(4, 0, 60, 5000, mode=2)
        soff = np.arange(61, dtype=np.uint64) * 5000
        w, off = nt.pack_chars(chars, soff, min_len=32)
        sk.submit(w, off)                                        # re-tiled + tails (general kernel)
        t, f1, p = sk.finish(counters=True, hist=True)           # apply, prefetch mode + histogram
        print("F1", [int(x) for x in f1], "nonzero counters", int((t != 0).sum()), "hist sum", int(p.sum()))
    os.environ["NTC_POOL_BLOCKS"] = "64"
    with nt.Sketch([32], rBits=20, sBits=7) as sk:              # pool exhaustion: direct increments
        sk.submit(words, None, n, stride)
        sk.submit(words, None, n, stride)
        _, f1, p = sk.finish(counters=False, hist=True)
        print("F1", [int(x) for x in f1], "hist sum", int(p.sum()))


if __name__ == "__main__":
    main()
