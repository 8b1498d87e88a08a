#!/usr/bin/env python
"""End-to-end command lines side by side: the unmodified reference `ntcard_ref -t<cores>` and this repo's
`bin/ntcard` on the same FASTQ files (generator of SURVEY 8d, split over as many files as host cores because the
reference parallelises over files only).  Prints both Runtime(sec) lines and checks the .hist files are identical.
    python tools/bench_cli.py [n_reads] [k]"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntcard_b200 as nt  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    L = 150
    cores = os.cpu_count() or 1
    nfiles = min(cores, 16)
    ref = os.path.join(ROOT, "oracle", "_ref", "ntcard_ref")
    ours = os.path.join(ROOT, "bin", "ntcard")
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        per = n // nfiles
        files = []
        t0 = time.time()
        qual = b"I" * L
        for f in range(nfiles):
            a = nt.gen_ascii(1, f * per, per, L, mode=1, U=max(per // 4, 1)).reshape(per, L)
            rec = np.empty((per, 2 * L + 8), dtype=np.uint8)     # "@r\n" seq "\n+\n" qual "\n"
            rec[:, 0:3] = np.frombuffer(b"@r\n", dtype=np.uint8)
            rec[:, 3:3 + L] = a
            rec[:, 3 + L:6 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
            rec[:, 6 + L:6 + 2 * L] = np.frombuffer(qual, dtype=np.uint8)
            rec[:, 6 + 2 * L] = 10
            rec = rec[:, :7 + 2 * L]
            p = os.path.join(td, f"part{f}.fq")
            rec.tofile(p)
            files.append(p)
        size = sum(os.path.getsize(p) for p in files)
        print(f"{nfiles} FASTQ files, {n} reads x {L} bp, {size / 1e9:.2f} GB, written in {time.time() - t0:.1f} s; host cores {cores}", flush=True)
        out = {}
        for tag, exe in (("reference", ref), ("ntcard_b200", ours)):
            if not os.path.exists(exe):
                print(tag, "binary missing:", exe)
                continue
            pref = os.path.join(td, tag)
            t0 = time.time()
            r = subprocess.run([exe, f"-t{nfiles}", f"-k{k}", "-c64", "-p", pref] + files, capture_output=True, text=True,
                               env=dict(os.environ, NTC_CLI_TIMING="1"))
            for l in r.stderr.splitlines():
                if "[timing]" in l:
                    print(l)
            wall = time.time() - t0
            rt = [l for l in r.stderr.splitlines() if l.startswith("Runtime")]
            print(f"{tag}: rc={r.returncode} wall {wall:.2f} s  {rt[0] if rt else r.stderr[-200:]}", flush=True)
            h = f"{pref}_k{k}.hist"
            out[tag] = open(h).read() if os.path.exists(h) else None
        if len(out) == 2:
            print("hist files identical:", out["reference"] is not None and out["reference"] == out["ntcard_b200"])
            print(out["ntcard_b200"][:60].replace("\n", " | "))


if __name__ == "__main__":
    main()
