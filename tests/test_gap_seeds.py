"""Gap seeds (-g N): the reference's stRead / stHashIterator / NTMSM64 path (ntcard.cpp:160-171, 407-413;
nthash.hpp:620-678; pinned in the reference by data/test-gap_k12.hist.good).  CPU: the oracle's restatement against
the golden vectors made from the unmodified reference (tests/golden/gap_cases.json, oracle/make_golden.py --only gap)
and against the reference itself where oracle/_ref exists.  GPU: the device path (general kernel with a second
rolling hash over the gap window) against the oracle and the reference CLI's .hist files."""
import os
import random
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden

CLI = os.path.join(ROOT, "bin", "ntcard")


def _reads(oracle, g):
    S, n, L, mode, U = g
    a = oracle.gen_reads(S, 0, n, L, mode, U)
    return [bytes(a[i * L:(i + 1) * L]) for i in range(n)]


def test_oracle_gap_hashes_match_reference_vectors(oracle):
    for c in load_golden("gap_cases.json")["hashes"]:
        h = oracle.st_hash_seq(c["seq"].encode(), c["k"], c["gap"])
        assert [f"{int(x):#018x}" for x in h] == c["h"], (c["k"], c["gap"])


def test_oracle_gap_sketches_match_reference_digests(oracle):
    gold = load_golden("gap_cases.json")
    for c in gold["sketches"]:
        reads = _reads(oracle, gold["gens"][c["name"]])
        sk, tot = oracle.sketch_reads_gap(reads, c["k"], c["gap"], c["rBits"], c["sBits"], nthreads=4)
        assert int(tot[0]) == c["F1"]
        rB = 1 << c["rBits"]
        for t in range(2):
            d = oracle.table_digest(np.ascontiguousarray(sk[t * rB:(t + 1) * rB]))
            assert (d["nnz"], d["sum"], d["max"], f"{d['digest']:#018x}") == (c["tables"][t]["nnz"], c["tables"][t]["sum"],
                                                                               c["tables"][t]["max"], c["tables"][t]["digest"])
        F0, f = oracle.compest(np.ascontiguousarray(sk), None, c["rBits"], c["sBits"], 64)
        assert F0 == c["est"]["F0"] and [float(x) for x in f[1:65]] == c["est"]["f"]


def test_oracle_gap_equals_reference_on_random_input(oracle, reference):
    rng = random.Random(77)
    reads = [bytes(rng.choice(b"ACGTacgtNu") for _ in range(rng.randint(0, 300))) for _ in range(2000)]
    for k, g in ((12, 2), (13, 3), (32, 30), (40, 8)):
        a, ta = oracle.sketch_reads_gap(reads, k, g, 12, 2, nthreads=2)
        b, tb = reference.sketch_reads_gap(reads, k, g, 12, 2, nthreads=2)
        assert np.array_equal(a, b) and np.array_equal(ta, tb)
    # gap 0 is plain ntRead
    a, ta = oracle.sketch_reads_gap(reads, 20, 0, 12, 2)
    b, tb = oracle.sketch_reads(reads, [20], 12, 2)
    assert np.array_equal(a, b) and np.array_equal(ta, tb)


@pytest.mark.gpu
def test_gpu_gap_seeds_match_oracle_and_golden(oracle):
    import ntcard_b200 as nt
    gold = load_golden("gap_cases.json")
    for c in gold["sketches"]:
        reads = _reads(oracle, gold["gens"][c["name"]])
        want, wf1 = oracle.sketch_reads_gap(reads, c["k"], c["gap"], c["rBits"], c["sBits"], nthreads=4)
        with nt.Sketch([c["k"]], rBits=c["rBits"], sBits=c["sBits"]) as sk:
            sk.set_gap(c["gap"])
            half = len(reads) // 2
            sk.submit_reads(reads[:half])                      # ragged batch
            L = len(reads[0])
            if all(len(r) == L for r in reads[half:]) and b"N" not in b"".join(reads[half:]):
                words = nt.pack_reads(reads[half:])[0]         # same records as one uniform-stride batch
                stride = len(words) // (len(reads) - half)
                sk.submit(words, None, len(reads) - half, stride)
            else:
                sk.submit_reads(reads[half:])
            t, f1, p = sk.finish(counters=True, hist=True)
            assert int(f1[0]) == c["F1"] == int(wf1[0])
            assert np.array_equal(t.reshape(-1), want)
            F0, f = nt.estimate(p_hist=p[0], rBits=c["rBits"], sBits=c["sBits"], covMax=64)
            assert F0 == c["est"]["F0"] and [float(x) for x in f[1:]] == c["est"]["f"]
            # back to plain ntHash on the same context
            sk.reset()
            sk.set_gap(0)
            sk.submit_reads(reads)
            t0, f10, _ = sk.finish(counters=True, hist=False)
            w0, wf0 = oracle.sketch_reads(reads, [c["k"]], c["rBits"], c["sBits"], nthreads=4)
            assert np.array_equal(t0.reshape(-1), w0) and np.array_equal(f10, wf0)
    with nt.Sketch([12, 32], rBits=12, sBits=7) as sk:
        with pytest.raises(nt.NtcError):
            sk.set_gap(2)                                       # one k only (ntcard.cpp:397)
    with nt.Sketch([12], rBits=12, sBits=7) as sk:
        with pytest.raises(nt.NtcError):
            sk.set_gap(3)                                       # parity (ntcard.cpp:382)
        with pytest.raises(nt.NtcError):
            sk.set_gap(12)


@pytest.mark.gpu
def test_cli_gap_matches_reference_cli(tmp_path, oracle):
    gold = load_golden("gap_cases.json")["cli"]
    g = gold["gen_a"]
    reads = _reads(oracle, (g["S"], g["n"], g["L"], g["mode"], g["U"]))
    fq = os.path.join(str(tmp_path), "a.fq")
    with open(fq, "w") as f:
        for i, r in enumerate(reads):
            f.write(f"@r{i}\n{r.decode()}\n+\n{'I' * len(r)}\n")
    pref = os.path.join(str(tmp_path), "out")
    for tag, args, k in (("fq_k12_g2_c50", ["-k12", "-g2", "-c50"], 12), ("fq_k32_g8_c20", ["-k32", "-g8", "-c20"], 32)):
        r = subprocess.run([CLI] + args + ["-p", pref, fq], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert open(f"{pref}_k{k}.hist").read() == gold[tag]
    for bad in (["-k12", "-g3"], ["-k12,32", "-g2"]):
        r = subprocess.run([CLI] + bad + ["-p", pref, fq], capture_output=True, text=True)
        assert r.returncode != 0
