#!/usr/bin/env python
"""The documented drop-in path: reads -> ntc_pack_seqs (N-split, 2-bit pack) -> RAGGED batches -> ntc_submit.

    python tools/bench_ragged.py [--reads 4000000] [--k 32]

150 bp reads with N runs (generator mode 2: 0..3 runs of 1..20 N per read) packed on the host into ragged batches of 500 k reads;
timed: the device kernels of one pass (CUDA events inside the library, ntc_kernel_time) and the wall time from the first
ntc_submit to the histogram on the host, pipeline (auto: batches padded to one stride on the device, mixed-length tiles) against
the general kernel (roll64).  Histograms must be identical.  One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ntcard_b200 as nt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--per", type=int, default=500_000, help="reads per host batch")
    ap.add_argument("--roll64-first", action="store_true", help="time the general kernel before the pipeline (order effects)")
    args = ap.parse_args()
    n, L, k, per = args.reads, 150, args.k, args.per
    batches = []
    n_rec = n_words = 0
    for lo in range(0, n, per):
        cnt = min(per, n - lo)
        chars = nt.gen_ascii(4, lo, cnt, L, mode=2)
        soff = np.arange(cnt + 1, dtype=np.uint64) * L
        w, off = nt.pack_chars(chars, soff, min_len=k)
        pw, po = nt.PinnedBuffer(len(w)), nt.PinnedBuffer(len(off))
        pw.array[:] = w
        po.array[:] = off
        batches.append((pw, po))
        n_rec += len(off) - 1
        n_words += len(w)
    out = {"workload": f"{n} reads x {L} bp with N runs -> {n_rec} ragged records, {n_words * 4 / 1e6:.0f} MB packed, k={k}, s=7, r=27"}
    hists = {}
    with nt.Sketch([k], rBits=27, sBits=7) as sk:
        modes = [("pipeline", nt.KERNEL_BITSLICE), ("roll64", nt.KERNEL_ROLL64)]
        for name, kern in (modes[::-1] if args.roll64_first else modes):
            sk.set_kernel(kern)

            host = {"submit_ms": 0.0, "finish_ms": 0.0}

            def step():
                sk.reset()
                t_a = time.perf_counter()
                for pw, po in batches:
                    sk.submit(pw.array, po.array)
                t_b = time.perf_counter()
                r = sk.finish(counters=False, hist=True)
                host["submit_ms"] += (t_b - t_a) * 1e3
                host["finish_ms"] += (time.perf_counter() - t_b) * 1e3
                return r
            for _ in range(4):        # every staging slot of the ring of 3 has seen every batch size: no cudaMalloc inside the timed passes
                step()
            sk.kernel_time()          # reading resets the accumulated kernel time
            host["submit_ms"] = host["finish_ms"] = 0.0
            t0 = time.perf_counter()
            for _ in range(args.steps):
                _, f1, p = step()
            wall = (time.perf_counter() - t0) * 1e3 / args.steps
            kms = sk.kernel_time()[0] / args.steps
            hists[name] = p
            out[name] = {"wall_ms_per_pass": wall, "kernel_ms_per_pass": kms, "kmers_per_s_kernels": int(f1[0]) / (kms * 1e-3),
                         "kmers_per_s_wall": int(f1[0]) / (wall * 1e-3), "host_in_submit_ms": host["submit_ms"] / args.steps,
                         "host_in_finish_ms": host["finish_ms"] / args.steps, "launches": sk.stats()["launches"]}
            out["F1"] = int(f1[0])
    assert np.array_equal(hists["pipeline"], hists["roll64"])
    out["histograms_identical"] = True
    print(json.dumps(out))


if __name__ == "__main__":
    main()
