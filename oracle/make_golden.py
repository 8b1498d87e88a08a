#!/usr/bin/env python
"""Generate tests/golden/*.json from the UNMODIFIED reference compiled here
(oracle/_ref, see oracle/Makefile).  Run in the build container only -- the GPU
box has no /root/reference; the committed fixtures travel instead.

    python oracle/make_golden.py

Every value below is produced by the reference's own code (ntHashIterator,
ntRead/ntComp, compEst, and the ntcard_ref CLI); the oracle restatement and the
CUDA path are both tested against these files.
"""
import json
import os
import random
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import HLL_REF_CLI, HllReference, Oracle, Reference, REF_CLI, build  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rand_seq(rng, L, alphabet=b"ACGT", junk=0.0, junk_chars=b"NnXRY-.*"):
    out = bytearray()
    for _ in range(L):
        if junk and rng.random() < junk:
            out.append(rng.choice(junk_chars))
        else:
            out.append(rng.choice(alphabet))
    return bytes(out)


def gap_golden():
    """tests/golden/gap_cases.json: the reference's gap-seed mode (-g; stRead / stHashIterator / NTMSM64,
    ntcard.cpp:160-171, 407-413; nthash.hpp:620-678) -- hash vectors, sketch digests, CLI output."""
    build(ref=True)
    orc, ref = Oracle(), Reference()

    def gen(S, n, L, mode=0, U=0):
        a = orc.gen_reads(S, 0, n, L, mode, U)
        return [bytes(a[i * L:(i + 1) * L]) for i in range(n)]

    out = {"source": "stRead / stHashIterator / NTMSM64 of the unmodified reference via oracle/_ref (ref_set_gap builds the seed the way "
                     "ntcard.cpp:407-413 does); reads from the SURVEY 8d generator", "hashes": [], "sketches": [], "cli": {}}
    seq = b"GAGTGTCAAACATTCAGACAACAGCAGGGGTGCTCTGGAATCCTATGTGAGGAACAAACATTCAGGCCACAGTAGNACGTTGCATGCCAGTNNACGATCGATCGGATCGATTACGAT"
    for k, g in ((12, 2), (20, 18), (31, 5), (32, 8), (33, 1), (64, 32)):
        h = ref.st_hash_seq(seq, k, g)
        out["hashes"].append({"seq": seq.decode(), "k": k, "gap": g, "h": [f"{int(x):#018x}" for x in h]})
    for name, reads, k, g, rBits, sBits in (("rep_k12_g2", gen(11, 4000, 150, 1, 500), 12, 2, 16, 3),
                                            ("rep_k32_g8", gen(12, 4000, 150, 1, 500), 32, 8, 18, 7),
                                            ("nmode_k31_g5", gen(4, 2000, 400, 2, 0), 31, 5, 18, 7),
                                            ("uni_k64_g32", gen(13, 3000, 150, 0, 0), 64, 32, 18, 7)):
        sk, tot = ref.sketch_reads_gap(reads, k, g, rBits, sBits, nthreads=4)
        rB = 1 << rBits
        tabs = []
        for t in range(2):
            d = orc.table_digest(np.ascontiguousarray(sk[t * rB:(t + 1) * rB]))
            d["digest"] = f"{d['digest']:#018x}"
            tabs.append(d)
        F0, fm = ref.compest(np.ascontiguousarray(sk), rBits, sBits)
        out["sketches"].append({"name": name, "k": k, "gap": g, "rBits": rBits, "sBits": sBits, "F1": int(tot[0]), "tables": tabs,
                                "est": {"F0": float(F0), "f": [float(x) for x in fm[1:65]]}})
    out["gens"] = {"rep_k12_g2": [11, 4000, 150, 1, 500], "rep_k32_g8": [12, 4000, 150, 1, 500], "nmode_k31_g5": [4, 2000, 400, 2, 0],
                   "uni_k64_g32": [13, 3000, 150, 0, 0]}
    with tempfile.TemporaryDirectory() as td:
        reads = gen(1, 20000, 150, 1, 2500)
        fq = os.path.join(td, "a.fq")
        with open(fq, "w") as f:
            for i, r in enumerate(reads):
                f.write(f"@r{i}\n{r.decode()}\n+\n{'I' * len(r)}\n")
        for tag, args in (("fq_k12_g2_c50", ["-k12", "-g2", "-c50"]), ("fq_k32_g8_c20", ["-k32", "-g8", "-c20"])):
            pref = os.path.join(td, "out")
            subprocess.run([REF_CLI] + args + ["-p", pref, fq], capture_output=True, text=True, check=True)
            k = args[0][2:]
            with open(os.path.join(td, f"out_k{k}.hist")) as f:
                out["cli"][tag] = f.read()
        out["cli"]["gen_a"] = {"S": 1, "n": 20000, "L": 150, "mode": 1, "U": 2500}
    with open(os.path.join(GOLD, "gap_cases.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("gap_cases.json", os.path.getsize(os.path.join(GOLD, "gap_cases.json")))


def fnv_bytes(a):
    """FNV-1a-64 over a uint8 array (offset 0xcbf29ce484222325, prime 0x100000001b3)."""
    h = 0xcbf29ce484222325
    for v in bytes(a):
        h = ((h ^ v) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def hll_golden():
    """tests/golden/hll_cases.json: nthll, the HyperLogLog estimator beside ntcard (nthll.cpp:92-104 ntComp / ntRead,
    :243-254 the estimate) -- register digests from the reference's own ntRead, and the unmodified CLI's output line."""
    build(ref=True)
    orc, ref = Oracle(), HllReference()

    def gen(S, n, L, mode=0, U=0):
        a = orc.gen_reads(S, 0, n, L, mode, U)
        return [bytes(a[i * L:(i + 1) * L]) for i in range(n)]

    out = {"source": "ntRead / ntComp of the unmodified reference nthll.cpp via oracle/_ref/libnthll_ref.so; the estimate is computed "
                     "by oracle/_ref/nthll_ref (the unmodified CLI) for the cli cases; reads from the SURVEY 8d generator "
                     "(S, n, L, mode, U)", "registers": [], "cli": {}}
    for name, g, k, nBits in (("rep_k32_b16", [21, 6000, 150, 1, 750], 32, 16), ("uni_k64_b16", [22, 6000, 150, 0, 0], 64, 16),
                              ("nmode_k31_b12", [4, 2000, 400, 2, 0], 31, 12), ("rep_k12_b4", [23, 3000, 150, 1, 300], 12, 4),
                              ("uni_k96_b18", [24, 4000, 150, 0, 0], 96, 18), ("uni_k150_b10", [25, 3000, 150, 0, 0], 150, 10),
                              ("long_k20_b17", [26, 40, 20000, 2, 0], 20, 17), ("uni_k1_b1", [27, 50, 40, 0, 0], 1, 1)):
        reads = gen(*g)
        regs = ref.hll_registers(reads, k, nBits, nthreads=4)
        case = {"name": name, "gen": g, "k": k, "nBits": nBits, "digest": f"{fnv_bytes(regs):#018x}", "max": int(regs.max()),
                "nonzero": int(np.count_nonzero(regs)), "sum": int(regs.astype(np.uint64).sum())}
        if nBits <= 4:
            case["regs"] = [int(x) for x in regs]
        out["registers"].append(case)
    with tempfile.TemporaryDirectory() as td:
        g = [1, 20000, 150, 1, 2500]
        reads = gen(*g)
        fq, fa, sam = (os.path.join(td, n) for n in ("a.fq", "a.fa", "a.sam"))
        with open(fq, "w") as f:
            for i, r in enumerate(reads):
                f.write(f"@r{i}\n{r.decode()}\n+\n{'I' * len(r)}\n")
        with open(fa, "w") as f:
            for i, r in enumerate(reads):
                s_ = r.decode()
                f.write(f">r{i}\n" + "\n".join(s_[j:j + 60] for j in range(0, len(s_), 60)) + "\n")
        with open(sam, "w") as f:
            f.write("@HD\tVN:1.6\n@SQ\tSN:x\tLN:1000\n")
            for i, r in enumerate(reads):
                f.write(f"r{i}\t4\t*\t0\t0\t*\t*\t0\t0\t{r.decode()}\t{'I' * len(r)}\n")
        for tag, args, files in (("fq_k32", ["-k32"], [fq]), ("fq_default", [], [fq]), ("fa_k32", ["-k32"], [fa]), ("sam_k32", ["-k32"], [sam]),
                                 ("fq_k12_b12_t2", ["-k12", "-b12", "-t2"], [fq, fa]), ("fq_k151", ["-k151"], [fq])):
            r = subprocess.run([HLL_REF_CLI] + args + files, capture_output=True, text=True, check=True)
            out["cli"][tag] = {"args": args, "files": [os.path.basename(x) for x in files], "stdout": r.stdout}
        out["cli"]["gen_a"] = {"S": g[0], "n": g[1], "L": g[2], "mode": g[3], "U": g[4]}
    with open(os.path.join(GOLD, "hll_cases.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("hll_cases.json", os.path.getsize(os.path.join(GOLD, "hll_cases.json")))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--only" and sys.argv[2] == "gap":
        return gap_golden()
    if len(sys.argv) > 2 and sys.argv[1] == "--only" and sys.argv[2] == "hll":
        return hll_golden()
    build(ref=True)
    os.makedirs(GOLD, exist_ok=True)
    ref = Reference()
    orc = Oracle()  # used ONLY for its read generator (ours, not the reference's)

    # ---- 1. ntHash vectors ---------------------------------------------------
    rng = random.Random(20261017)
    seqs = [
        b"ACGTACACTGGACTGAGTCT",
        b"AGACTCAGTCCAGTGTACGT",
        b"GAGTGTCAAACATTCAGACAACAGCAGGGGTGCTCTGGAATCCTATGTGAGGAACAAACATTCAGGCCACAGTAG",
        b"ACGTNACGTACGTTTGA",
        b"acgtacactggactgagtct",
        b"ACGUACACUGGACUGAGUCU",
        b"NNNNNNNNNN",
        b"ACGT",
        b"",
        b"ACGTACGTACGTNACGTACGTACGTNNACGTACGTACGTACGTACGTACGT",
    ]
    for L in (40, 150, 151, 300, 1000):
        seqs.append(rand_seq(rng, L))
        seqs.append(rand_seq(rng, L, alphabet=b"ACGTacgtUu", junk=0.03))
    vec = []
    for s in seqs:
        for k in (1, 4, 12, 18, 20, 31, 32, 33, 64, 70, 96, 128):
            h, p = ref.hash_seq(s, k)
            x = 0
            for v in h:
                x ^= int(v)
            vec.append({
                "seq": s.decode("latin1"), "k": k, "n": int(len(h)),
                "pos_first": int(p[0]) if len(p) else -1, "pos_last": int(p[-1]) if len(p) else -1,
                "pos_sum": int(np.sum(p.astype(np.uint64))),
                "first": f"{int(h[0]):#018x}" if len(h) else None,
                "last": f"{int(h[-1]):#018x}" if len(h) else None,
                "xor": f"{x:#018x}",
                "hashes": [f"{int(v):#018x}" for v in h] if len(s) <= 80 else None,
                "pos": [int(v) for v in p] if len(s) <= 80 else None,
            })
    kmer = b"ACGTACACTGGACTGAGTCT"
    fh, rh = ref.kmer_hashes(kmer)
    mst = {str(k): {chr(c): f"{ref.mstab(c, k):#018x}" for c in b"ACGT"} for k in
           (1, 2, 12, 30, 31, 32, 33, 34, 62, 63, 64, 65, 66, 96, 100, 127, 128, 200, 250, 1023, 1024)}
    with open(os.path.join(GOLD, "nthash_vectors.json"), "w") as f:
        json.dump({"source": "vendor/ntHash ntHashIterator / NTF64 / NTR64 / msTab via oracle/_ref",
                   "unit_test_kmer": {"seq": kmer.decode(), "fh": f"{fh:#018x}", "rh": f"{rh:#018x}"},
                   "srol_k_of_seed": mst, "vectors": vec}, f, indent=0)

    # ---- 2. sketches (ntRead/ntComp) on seeded inputs ------------------------
    sk_cases = []

    def sketch_case(name, reads, kList, rBits, sBits, keep_full=False):
        sk, tot = ref.sketch_reads(reads, kList, rBits, sBits, nthreads=4)
        rB = 1 << rBits
        tabs = []
        for ki in range(len(kList)):
            for t in range(2):
                tab = sk[(ki * 2 + t) * rB:(ki * 2 + t + 1) * rB]
                d = orc.table_digest(np.ascontiguousarray(tab))
                d["digest"] = f"{d['digest']:#018x}"
                if d["nnz"] == 0:
                    d["first"] = -1
                if keep_full:
                    nz = np.nonzero(tab)[0]
                    d["nonzero"] = [[int(i), int(tab[i])] for i in nz]
                tabs.append(d)
        # estimator on the same sketch (full recurrence), rows 1..64 kept
        est = []
        for ki in range(len(kList)):
            F0, fm = ref.compest(np.ascontiguousarray(sk[ki * 2 * rB:(ki + 1) * 2 * rB]), rBits, sBits)
            est.append({"F0": float(F0), "f": [float(x) for x in fm[1:65]]})
        return {"name": name, "k": list(kList), "rBits": rBits, "sBits": sBits,
                "F1": [int(x) for x in tot], "tables": tabs, "est": est}

    def gen(S, n, L, mode=0, U=0):
        a = orc.gen_reads(S, 0, n, L, mode, U)
        return [bytes(a[i * L:(i + 1) * L]) for i in range(n)]

    # FX1 of SURVEY 8c: seed 1, L 150, n 200000, U 25000 (repeat mode), s=7, r=27
    fx1 = gen(1, 200000, 150, 1, 25000)
    c = sketch_case("FX1", fx1, [12, 31, 32, 64], 27, 7)
    c["gen"] = {"S": 1, "n": 200000, "L": 150, "mode": 1, "U": 25000}
    sk_cases.append(c)
    # small table, many collisions; s=3 so that plenty is sampled
    c = sketch_case("small_r12_s3", gen(7, 2000, 100, 1, 500), [12, 32], 12, 3, keep_full=False)
    c["gen"] = {"S": 7, "n": 2000, "L": 100, "mode": 1, "U": 500}
    sk_cases.append(c)
    # Ns (mode 2), s=11 / s=7
    for s in (7, 11):
        c = sketch_case(f"nmode_s{s}", gen(4, 3000, 400, 2, 0), [31, 64], 20, s)
        c["gen"] = {"S": 4, "n": 3000, "L": 400, "mode": 2, "U": 0}
        sk_cases.append(c)
    # counter wrap: one k-mer-rich read repeated 70000 times must wrap mod 65536
    # (uint16 ++, ntcard.cpp:133,143).  s=1 so everything with top bit 0 is sampled.
    wrap_read = gen(9, 1, 40, 0, 0)[0]
    c = sketch_case("wrap_u16", [wrap_read] * 70000, [32], 10, 1, keep_full=True)
    c["gen"] = {"S": 9, "n": 1, "L": 40, "mode": 0, "U": 0, "repeat": 70000}
    sk_cases.append(c)
    # ragged lengths incl. shorter than k and empty
    rr = random.Random(5)
    ragged = [rand_seq(rr, rr.choice((0, 1, 5, 11, 12, 13, 31, 32, 33, 63, 64, 65, 100, 150, 250, 1000)),
                       alphabet=b"ACGT", junk=0.01) for _ in range(1500)]
    c = sketch_case("ragged_r16_s2", ragged, [12, 32, 64], 16, 2)
    c["reads"] = [r.decode("latin1") for r in ragged]
    sk_cases.append(c)
    with open(os.path.join(GOLD, "sketch_cases.json"), "w") as f:
        json.dump({"source": "ntRead/ntComp/compEst of ntcard.cpp via oracle/_ref; reads from the SURVEY 8d generator",
                   "digest": "FNV-1a-64 over (idx,count) of nonzero buckets in index order",
                   "cases": sk_cases}, f, indent=0)

    # ---- 3. compEst on synthetic counter-value histograms ---------------------
    est_cases = []
    rs = np.random.RandomState(3)
    for rBits, sBits, lam in ((20, 7, 0.02), (16, 3, 0.3), (18, 11, 1.5)):
        rB = 1 << rBits
        sk = np.minimum(rs.poisson(lam, size=2 * rB) * rs.randint(1, 9, size=2 * rB), 65535).astype(np.uint16)
        F0, fm = ref.compest(sk, rBits, sBits)
        p = np.stack([np.bincount(sk[:rB], minlength=65536), np.bincount(sk[rB:], minlength=65536)]).astype(np.uint32)
        nzp = [[int(t), int(i), int(p[t, i])] for t in range(2) for i in np.nonzero(p[t])[0]]
        est_cases.append({"rBits": rBits, "sBits": sBits, "p_hist_nonzero": nzp, "F0": float(F0),
                          "f": [float(x) for x in fm[1:1001]]})
    # empty sketch early-out (ntcard.cpp:262-264)
    sk = np.zeros(2 << 12, dtype=np.uint16)
    F0, fm = ref.compest(sk, 12, 7)
    est_cases.append({"rBits": 12, "sBits": 7, "p_hist_nonzero": [[0, 0, 4096], [1, 0, 4096]], "F0": float(F0),
                      "f": [float(x) for x in fm[1:1001]]})
    with open(os.path.join(GOLD, "compest_cases.json"), "w") as f:
        json.dump({"source": "compEst of ntcard.cpp:237-275 via oracle/_ref", "cases": est_cases}, f, indent=0)

    # ---- 4. the unmodified CLI on small files (formats + .hist layout) --------
    cli = {}
    with tempfile.TemporaryDirectory() as td:
        reads = gen(1, 20000, 150, 1, 2500)
        nreads = gen(4, 2000, 400, 2, 0)
        fq = os.path.join(td, "a.fq")
        with open(fq, "w") as f:
            for i, r in enumerate(reads):
                f.write(f"@r{i}\n{r.decode()}\n+\n{'I' * len(r)}\n")
        fa = os.path.join(td, "a.fa")
        with open(fa, "w") as f:
            for i, r in enumerate(reads):
                s = r.decode()
                f.write(f">r{i}\n" + "\n".join(s[j:j + 60] for j in range(0, len(s), 60)) + "\n")
        sam = os.path.join(td, "a.sam")
        with open(sam, "w") as f:
            f.write("@HD\tVN:1.6\n@SQ\tSN:x\tLN:100\n")
            for i, r in enumerate(reads):
                f.write(f"r{i}\t4\t*\t0\t0\t*\t*\t0\t0\t{r.decode()}\t{'I' * len(r)}\n")
        fqn = os.path.join(td, "n.fq")
        with open(fqn, "w") as f:
            for i, r in enumerate(nreads):
                f.write(f"@r{i}\n{r.decode()}\n+\n{'I' * len(r)}\n")

        def run(args, files):
            pref = os.path.join(td, "out")
            p = subprocess.run([REF_CLI] + args + ["-p", pref] + files, capture_output=True, text=True, check=True)
            outs = {}
            for fn in sorted(os.listdir(td)):
                if fn.startswith("out_k") and fn.endswith(".hist"):
                    with open(os.path.join(td, fn)) as f:
                        outs[fn[4:]] = f.read()
                    os.remove(os.path.join(td, fn))
            return outs

        cli["gen_a"] = {"S": 1, "n": 20000, "L": 150, "mode": 1, "U": 2500}
        cli["gen_n"] = {"S": 4, "n": 2000, "L": 400, "mode": 2, "U": 0}
        cli["fq_k12_k32_c50"] = run(["-k12,32", "-c50"], [fq])
        cli["fa_k12_k32_c50"] = run(["-k12,32", "-c50"], [fa])
        cli["sam_k12_k32_c50"] = run(["-k12,32", "-c50"], [sam])
        cli["nfq_k31_c20"] = run(["-k31", "-c20"], [fqn])
        cli["two_files_t2_k64_c20"] = run(["-k64", "-c20", "-t2"], [fq, fqn])
        # compact output (-o): file + stderr lines
        o = os.path.join(td, "compact.tsv")
        p = subprocess.run([REF_CLI, "-k12,64", "-c8", "-o", o, fq], capture_output=True, text=True, check=True)
        with open(o) as f:
            cli["compact_k12_k64_c8"] = {"file": f.read(),
                                         "stderr": "".join(l + "\n" for l in p.stderr.splitlines() if not l.startswith("Runtime"))}
    with open(os.path.join(GOLD, "cli_cases.json"), "w") as f:
        json.dump({"source": "unmodified ntcard CLI (oracle/_ref/ntcard_ref); inputs written by this script from the SURVEY 8d generator",
                   "cases": cli}, f, indent=0)
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)))


if __name__ == "__main__":
    main()
