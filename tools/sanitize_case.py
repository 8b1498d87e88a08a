#!/usr/bin/env python
"""A small pass through every kernel of the pipeline (scan, hit staged + plain, apply zero / prefetch mode, fallback,
re-tiling, general kernel, histograms; round 2: headerless batches, device-resident ragged batches, the peer-memory reduction
between two contexts, the nthll pre-filter, the fused kernel) for compute-sanitizer:
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


def round2():
    import torch
    n, L = 3000, 150
    stride = nt.stride_words(L)
    words = nt.gen_packed(5, 0, n, L, 1, n // 4, stride)
    del os.environ["NTC_POOL_BLOCKS"]
    # headerless uniform batch, device-resident ragged batch (check_offsets_kernel, pad_ragged)
    with nt.Sketch([25, 32], rBits=20, sBits=7) as sk:
        wpr = (L + 15) // 16
        sk.submit_bases(np.ascontiguousarray(words.reshape(n, stride)[:, 1:1 + wpr]).reshape(-1), n, L)
        chars = nt.gen_ascii(4, 0, 2500, 150, mode=2)
        w, off = nt.pack_chars(chars, np.arange(2501, dtype=np.uint64) * 150, min_len=25)
        dw, do = torch.from_numpy(w.view(np.int32)).cuda(), torch.from_numpy(off.view(np.int32)).cuda()
        sk.submit_device(dw.data_ptr(), len(w), len(off) - 1, 0, d_off=do.data_ptr())
        _, f1, p = sk.finish(counters=False, hist=True)
        print("round 2 headerless + device ragged: F1", [int(x) for x in f1], "hist sum", int(p.sum()))
    # peer-memory reduction between two contexts (apply_owned_kernel, log_status_kernel, hist_slices)
    ranks = [nt.Sketch([32], rBits=20, sBits=7) for _ in range(2)]
    for r, sk in enumerate(ranks):
        sk.peer_attach_contexts(r, ranks)
    status = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in ranks]
    hist = [torch.zeros(2 * 65536, dtype=torch.int32, device="cuda") for _ in ranks]
    for r, sk in enumerate(ranks):
        sk.submit(words[r * 1500 * stride:(r + 1) * 1500 * stride], None, 1500, stride)
        sk.log_status_device(status[r].data_ptr())
    for sk in ranks:
        sk.stream_sync()
    tot = torch.stack(status).sum(dim=0)
    for r, sk in enumerate(ranks):
        sk.reduce_owned(tot.data_ptr(), hist[r].data_ptr())
    for sk in ranks:
        sk.stream_sync()
    print("round 2 peer reduction: hist sum", int(torch.stack(hist).sum()))
    for sk in ranks:
        sk.close()
    # nthll pre-filter (scan MODE 2 / 1 + hll_hit_kernel + hll_min_kernel): 1024 registers, so that the first batch (general kernel)
    # lifts every register past 4 and the next ones take the filter at the level the device reads, then the fixed 13-bit one
    with nt.HllSketch(32, 10) as h:
        big = nt.gen_packed(7, 0, 5000, L, 0, 0, stride)
        h.submit(big[:4096 * stride], None, 4096, stride)   # general kernel + hll_min
        l0 = h.stats()["launches"]
        h.submit(big, None, 5000, stride)                    # pre-filter path, level from the device word
        l1 = h.stats()["launches"]
        for _ in range(12):
            h.submit(big, None, 5000, stride)
        r_, nk = h.finish()
        print("round 2 nthll: k-mers", nk, "min / max register", int(r_.min()), int(r_.max()), "launches of the 2nd batch", l1 - l0)
    os.environ["NTC_FUSED"] = "1"
    with nt.Sketch([32, 64], rBits=20, sBits=7) as sk:      # fused scan + hash + append kernel
        sk.submit(words, None, n, stride)
        _, f1, p = sk.finish(counters=False, hist=True)
        print("round 2 fused kernel: F1", [int(x) for x in f1], "hist sum", int(p.sum()))
    del os.environ["NTC_FUSED"]


if __name__ == "__main__":
    if "--round2-only" not in sys.argv:
        main()
    else:
        os.environ["NTC_POOL_BLOCKS"] = "64"
    round2()
