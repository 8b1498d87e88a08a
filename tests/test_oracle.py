"""Pins oracle/ntcard_oracle.c: reference unit-test vectors, golden fixtures
generated from the reference (oracle/make_golden.py), and -- where oracle/_ref
exists -- the reference itself on random inputs.  CPU only."""
import random

import numpy as np
import pytest

from conftest import load_golden


def H(x):
    return int(x, 16)


# ---- vendor/ntHash/unittest/UnitTests.cpp ------------------------------------
KMER = b"ACGTACACTGGACTGAGTCT"


def test_unit_invariant_hash(oracle):
    # UnitTests.cpp:43-54: element 0 of the h=3 vector is the canonical hash ntCard uses
    h, p = oracle.hash_seq(KMER, 20)
    assert list(h) == [10434435546371013747] and list(p) == [0]


def test_unit_reverse_complement(oracle):
    # UnitTests.cpp:56-68
    h1, _ = oracle.hash_seq(KMER, 20)
    h2, _ = oracle.hash_seq(b"AGACTCAGTCCAGTGTACGT", 20)
    assert h1[0] == h2[0]


def test_unit_rolling_equals_base(oracle):
    # UnitTests.cpp:70-93
    roll, _ = oracle.hash_seq(KMER, 18)
    for i, km in enumerate((b"ACGTACACTGGACTGAGT", b"CGTACACTGGACTGAGTC", b"GTACACTGGACTGAGTCT")):
        base, _ = oracle.hash_seq(km, 18)
        assert roll[i] == base[0]
    assert [int(x) for x in roll] == [0xb6af114878e14baf, 0x18dd0aef1339c397, 0x4b4ccbf8e3b3ae2b]


def test_unit_rna_equals_dna(oracle):
    # UnitTests.cpp:121-132 (as intended: U hashes like T)
    h1, _ = oracle.hash_seq(KMER, 20)
    h2, _ = oracle.hash_seq(b"ACGUACACUGGACUGAGUCU", 20)
    assert h1[0] == h2[0]


# ---- golden vectors from the reference ---------------------------------------
def test_golden_nthash_vectors(oracle):
    g = load_golden("nthash_vectors.json")
    fh, rh = oracle.kmer_hashes(g["unit_test_kmer"]["seq"].encode())
    assert fh == H(g["unit_test_kmer"]["fh"]) and rh == H(g["unit_test_kmer"]["rh"])
    seeds = {c: oracle.seed(ord(c)) for c in "ACGT"}
    for k, row in g["srol_k_of_seed"].items():
        for c, v in row.items():
            assert oracle.srol_n(seeds[c], int(k)) == H(v)
    for v in g["vectors"]:
        s = v["seq"].encode("latin1")
        h, p = oracle.hash_seq(s, v["k"])
        assert len(h) == v["n"], (v["seq"][:30], v["k"])
        if v["n"]:
            x = 0
            for t in h:
                x ^= int(t)
            assert int(h[0]) == H(v["first"]) and int(h[-1]) == H(v["last"]) and x == H(v["xor"])
            assert int(p[0]) == v["pos_first"] and int(p[-1]) == v["pos_last"]
            assert int(p.astype(np.uint64).sum()) == v["pos_sum"]
        if v["hashes"] is not None:
            assert [int(t) for t in h] == [H(t) for t in v["hashes"]]
            assert [int(t) for t in p] == v["pos"]


def _reads_of(case, oracle):
    if "reads" in case:
        return [r.encode("latin1") for r in case["reads"]]
    g = case["gen"]
    a = oracle.gen_reads(g["S"], 0, g["n"], g["L"], g["mode"], g["U"])
    reads = [bytes(a[i * g["L"]:(i + 1) * g["L"]]) for i in range(g["n"])]
    return reads * g.get("repeat", 1)


@pytest.mark.parametrize("name", ["FX1", "small_r12_s3", "nmode_s7", "nmode_s11", "wrap_u16", "ragged_r16_s2"])
def test_golden_sketch_cases(oracle, name):
    case = next(c for c in load_golden("sketch_cases.json")["cases"] if c["name"] == name)
    reads = _reads_of(case, oracle)
    rB = 1 << case["rBits"]
    sk, tot = oracle.sketch_reads(reads, case["k"], case["rBits"], case["sBits"], nthreads=4)
    assert [int(x) for x in tot] == case["F1"]
    for i, gt in enumerate(case["tables"]):
        tab = np.ascontiguousarray(sk[i * rB:(i + 1) * rB])
        d = oracle.table_digest(tab)
        assert d["nnz"] == gt["nnz"] and d["sum"] == gt["sum"] and d["max"] == gt["max"]
        assert d["digest"] == H(gt["digest"])
        if "nonzero" in gt:
            assert [[int(j), int(tab[j])] for j in np.nonzero(tab)[0]] == gt["nonzero"]
    for ki, ge in enumerate(case["est"]):
        F0, f = oracle.compest(np.ascontiguousarray(sk[ki * 2 * rB:(ki + 1) * 2 * rB]), None,
                               case["rBits"], case["sBits"], 64)
        assert F0 == ge["F0"] and [float(x) for x in f[1:65]] == ge["f"]


def test_fx1_survey_numbers(oracle):
    """the FX1 numbers recorded from the reference CLI."""
    case = next(c for c in load_golden("sketch_cases.json")["cases"] if c["name"] == "FX1")
    i32 = case["k"].index(32)
    assert case["F1"][i32] == 23800000 and case["est"][i32]["F0"] == 2953021.0
    assert case["est"][i32]["f"][7] == 2952826.0
    t0, t1 = case["tables"][2 * i32], case["tables"][2 * i32 + 1]
    assert (t0["nnz"], t0["sum"], t0["max"], t0["first"], H(t0["digest"])) == (22944, 183576, 16, 6617, 0x842c569f4f284af1)
    assert (t1["nnz"], t1["sum"], t1["max"], t1["first"], H(t1["digest"])) == (23193, 185576, 16, 801, 0x0474fcb4d5dd2306)


def test_golden_compest(oracle):
    for c in load_golden("compest_cases.json")["cases"]:
        p = np.zeros((2, 65536), dtype=np.uint32)
        for t, i, v in c["p_hist_nonzero"]:
            p[t, i] = v
        F0, f = oracle.compest(None, p, c["rBits"], c["sBits"], 1000)
        assert F0 == c["F0"]
        assert [float(x) for x in f[1:1001]] == c["f"]


def test_compest_truncation_is_exact(oracle):
    """f[i] depends only on f[j<i] (ntcard.cpp:266-272): truncating at covMax changes nothing."""
    rs = np.random.RandomState(11)
    sk = (rs.poisson(0.05, size=2 << 14) * rs.randint(1, 5, size=2 << 14)).astype(np.uint16)
    F0a, fa = oracle.compest(sk, None, 14, 7, 3000)
    F0b, fb = oracle.compest(sk, None, 14, 7, 50)
    assert F0a == F0b and np.array_equal(fa[:51], fb[:51]) and not fb[51:].any()


def test_valid_window_definition(oracle):
    """ntHashIterator yields exactly the windows whose k chars are all in ACGTUacgtu
    (ntHashIterator.hpp:66-67,80-83) -- the rule the host packer relies on."""
    rng = random.Random(3)
    ok = set(b"ACGTUacgtu")
    for _ in range(200):
        L = rng.randint(0, 120)
        s = bytes(rng.choice(b"ACGTacgtUuNnX-") for _ in range(L))
        for k in (1, 3, 12, 31):
            _, p = oracle.hash_seq(s, k)
            want = [i for i in range(0, L - k + 1) if all(c in ok for c in s[i:i + k])]
            assert list(p) == want


def test_segments_reproduce_read(oracle):
    """Splitting a read at invalid chars and hashing the segments separately gives the
    same multiset of hashes and the same F1 (the host packer's N-split)."""
    rng = random.Random(4)
    ok = set(b"ACGTUacgtu")
    for _ in range(100):
        s = bytes(rng.choice(b"ACGTACGTACGTNn") for _ in range(rng.randint(1, 300)))
        for k in (4, 12, 32):
            h, _ = oracle.hash_seq(s, k)
            segs, cur = [], bytearray()
            for c in s:
                if c in ok:
                    cur.append(c)
                else:
                    segs.append(bytes(cur)); cur = bytearray()
            segs.append(bytes(cur))
            hs = [oracle.hash_seq(g, k)[0] for g in segs]
            h2 = np.concatenate(hs) if hs else np.zeros(0, dtype=np.uint64)
            assert np.array_equal(h, h2)


def test_sampling_rule(oracle):
    # ntcard.cpp:136-139; s=7: top byte 0x01 -> table 0, 0x7e/0x7f -> table 1
    assert oracle.sample_table(0x01 << 56, 7) == 0
    assert oracle.sample_table(0x7e << 56, 7) == 1 and oracle.sample_table(0x7f << 56 | 5, 7) == 1
    assert oracle.sample_table(0x02 << 56, 7) == 2 and oracle.sample_table(0, 7) == 2
    assert oracle.sample_table(1 << 52, 11) == 0 and oracle.sample_table(1023 << 53, 11) == 1
    # s=1: top bits 01 satisfy both tests; the later one (table 1) wins
    assert oracle.sample_table(1 << 62, 1) == 1 and oracle.sample_table(1 << 63, 1) == 2


# ---- the reference itself, where it was compiled ------------------------------
def test_oracle_vs_reference_random(oracle, reference):
    rng = random.Random(99)
    for c in range(256):
        assert oracle.seed(c) == reference.seed(c)
    for _ in range(200):
        v = rng.getrandbits(64)
        assert oracle.srol(v) == reference.srol(v) and oracle.sror(v) == reference.sror(v)
        assert oracle.sror(oracle.srol(v)) == v
    for _ in range(150):
        L = rng.randint(0, 400)
        s = bytes(rng.choice(b"ACGTacgtNUXu\x01\n") if rng.random() < 0.05 else rng.choice(b"ACGT") for _ in range(L))
        for k in (1, 12, 31, 32, 33, 64, 128):
            h1, p1 = oracle.hash_seq(s, k)
            h2, p2 = reference.hash_seq(s, k)
            assert np.array_equal(h1, h2) and np.array_equal(p1, p2)


def test_oracle_vs_reference_sketch(oracle, reference):
    a = oracle.gen_reads(21, 0, 5000, 150, 2, 0)
    reads = [bytes(a[i * 150:(i + 1) * 150]) for i in range(5000)]
    for s in (3, 7, 11):
        sk1, t1 = oracle.sketch_reads(reads, [12, 32, 96], 16, s, nthreads=3)
        sk2, t2 = reference.sketch_reads(reads, [12, 32, 96], 16, s, nthreads=3)
        assert np.array_equal(sk1, sk2) and np.array_equal(t1, t2)
