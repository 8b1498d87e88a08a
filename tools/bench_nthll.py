#!/usr/bin/env python
"""nthll on the device vs the reference's own ntRead / ntComp on the host cores.

    python tools/bench_nthll.py [--reads 10000000] [--k 32,64] [--bits 16] [--steps 5]

Per k: hll_kernel over device-resident packed reads (CUDA events around the launches, registers reset every step),
the same through host pinned buffers (H2D inside the timed region, registers read back, estimate on the host), and a
bounded sample of the workload through oracle/_ref/libnthll_ref.so with OpenMP over reads (when it was built).
Registers of the device run are checked against the oracle's on a 200 k-read prefix.  One JSON line per k."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ntcard_b200 as nt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--len", type=int, default=150)
    ap.add_argument("--k", default="32,64")
    ap.add_argument("--bits", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu-reads", type=int, default=2_000_000)
    args = ap.parse_args()
    n, L = args.reads, args.len
    stride = nt.stride_words(L)
    st = torch.cuda.Stream()
    d_words = torch.empty(n * stride, dtype=torch.int32, device="cuda")
    from oracle.pyoracle import HllReference, Oracle  # the checker and the CPU arm only
    orc = Oracle()
    for k in [int(x) for x in args.k.split(",")]:
        with torch.cuda.stream(st), nt.HllSketch(k, args.bits, stream=st.cuda_stream) as h:
            h.gen_packed_device(1, 0, n, L, 0, 0, stride, d_words.data_ptr())
            # parity on a prefix
            npre = min(n, 200_000)
            h.submit_device(d_words.data_ptr(), npre * stride, npre, stride)
            regs, nk = h.finish()
            a = orc.gen_reads(1, 0, npre, L, 0, 0)
            want = orc.hll_registers([bytes(a[i * L:(i + 1) * L]) for i in range(npre)], k, args.bits, nthreads=os.cpu_count())
            assert np.array_equal(regs, want) and nk == npre * (L - k + 1), "device registers differ from the oracle"

            def step():
                h.reset()
                h.submit_device(d_words.data_ptr(), n * stride, n, stride)
            for _ in range(3):
                step()
            h.sync()
            h.kernel_time()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(args.steps):
                step()
            e1.record(st)
            h.sync()
            ms = e0.elapsed_time(e1) / args.steps
            regs, nk = h.finish()
            kmers = n * (L - k + 1)
            assert nk == kmers
            est = nt.hll_estimate(regs, args.bits)
            # host buffers, 16 batches, H2D inside the timed region
            estride = nt.stride_words(L, align4=False)
            per = (n + 15) // 16
            pins = []
            for b in range(16):
                lo, hi = b * per, min(n, (b + 1) * per)
                pb = nt.PinnedBuffer((hi - lo) * estride)
                nt.gen_packed(1, lo, hi - lo, L, 0, 0, estride, out=pb.array)
                pins.append((pb, hi - lo))

            def step_e2e():
                h.reset()
                for pb, cnt in pins:
                    h.submit(pb.array, None, cnt, estride)
                r, _ = h.finish()
                return nt.hll_estimate(r, args.bits)
            for _ in range(2):
                step_e2e()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2 = step_e2e()
            e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            assert e2 == est
            for pb, _ in pins:
                pb.free()
        line = {"metric": "k-mers hashed/sec, nthll (HyperLogLog) on %d bp reads" % L, "k": k, "nBits": args.bits, "reads": n,
                "value": kmers / (ms * 1e-3), "ms_per_step": ms, "e2e": {"value": kmers / (e2e_ms * 1e-3), "ms_per_step": e2e_ms,
                                                                          "h2d_bytes_per_step": n * estride * 4, "d2h_bytes_per_step": 1 << args.bits},
                "estimate": int(est), "distinct_true": kmers, "algorithmic_GBps": n * (4 + (L + 3) // 4) / (ms * 1e-3) / 1e9}
        if HllReference.available():
            ref = HllReference()
            nc = min(n, args.cpu_reads)
            a = orc.gen_reads(1, 0, nc, L, 0, 0)
            reads = [bytes(a[i * L:(i + 1) * L]) for i in range(nc)]
            cores = os.cpu_count()
            ref.hll_registers(reads[:10000], k, args.bits, nthreads=cores)
            t0 = time.perf_counter()
            ref.hll_registers(reads, k, args.bits, nthreads=cores)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"kind": "reference", "cores": cores, "value": nc * (L - k + 1) / dt,
                                    "sample": f"{nc} reads in RAM through nthll.cpp's ntRead, OpenMP over reads"}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
