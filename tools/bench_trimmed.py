#!/usr/bin/env python
"""Trimmed reads: every record has its own length (100..150 bases, one stride) -- the mixed-length tiles of the scan kernel.

    python tools/bench_trimmed.py [--reads 10000000] [--k 32] [--steps 5]

Device-resident batch; pipeline (auto) against the general kernel (roll64): ms per pass (reset -> complete sketch), k-mers/s,
and the two counter-value histograms must be identical.  One JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ntcard_b200 as nt  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--sbits", type=int, default=7)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--lo", type=int, default=100)
    ap.add_argument("--hi", type=int, default=150)
    args = ap.parse_args()
    n, k, L = args.reads, args.k, args.hi
    stride = nt.stride_words(L)
    st = torch.cuda.Stream()
    out = {"workload": f"{n} reads, lengths uniform in [{args.lo}, {args.hi}], stride {stride} words, k={k}, s={args.sbits}, r=27"}
    with torch.cuda.stream(st):
        d_words = torch.empty(n * stride, dtype=torch.int32, device="cuda")
        with nt.Sketch([k], rBits=27, sBits=args.sbits, stream=st.cuda_stream) as sk:
            sk.gen_packed_device(1, 0, n, L, 0, 0, stride, d_words.data_ptr())
            sk.sync()
            g = torch.Generator(device="cuda")
            g.manual_seed(7)
            lens = torch.randint(args.lo, args.hi + 1, (n,), generator=g, device="cuda", dtype=torch.int32)
            d_words.view(n, stride)[:, 0] = lens          # bases past a record's new end stay: garbage padding
            kmers = int((lens.to(torch.int64) - k + 1).clamp(min=0).sum().item())
            st.synchronize()
            hists = {}
            for name, kern in (("pipeline", nt.KERNEL_AUTO), ("roll64", nt.KERNEL_ROLL64)):
                sk.set_kernel(kern)

                def step():
                    sk.reset()
                    sk.submit_device(d_words.data_ptr(), n * stride, n, stride)
                    sk.flush()
                for _ in range(3):
                    step()
                sk.sync()
                sk.stage_times()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(args.steps):
                    step()
                e1.record(st)
                sk.sync()
                ms = e0.elapsed_time(e1) / args.steps
                stages = [x / args.steps for x in sk.stage_times()]
                _, f1, p = sk.finish(counters=False, hist=True)
                assert int(f1[0]) == kmers, (int(f1[0]), kmers)
                hists[name] = p
                out[name] = {"ms_per_step": ms, "kmers_per_s": kmers / (ms * 1e-3), "stages_ms": {"scan": stages[0], "hit": stages[1], "apply": stages[2]}}
            assert np.array_equal(hists["pipeline"], hists["roll64"]), "pipeline and general kernel disagree"
            out["F1"] = kmers
            out["histograms_identical"] = True
    print(json.dumps(out))


if __name__ == "__main__":
    main()
