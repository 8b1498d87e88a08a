"""CPU-only checks of the product's host side: the C-ABI library loads and exports every symbol the
header declares, refuses to compute without a device (no CPU fallback), and its host helpers
(packer, generator, estimator) agree with the oracle / golden vectors."""
import ctypes
import os
import re

import numpy as np
import pytest

import ntcard_b200 as nt
from conftest import ROOT, load_golden


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "ntcard_b200.h")).read()
    declared = set(re.findall(r"\b(ntc_[a-z0-9_]+)\s*\(", hdr)) - {"ntc_ctx"}
    assert declared == set(nt.SYMBOLS), declared ^ set(nt.SYMBOLS)
    raw = ctypes.CDLL(nt.LIB_PATH)
    for s in declared:
        assert hasattr(raw, s), s


def test_no_cpu_fallback():
    if nt.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nt.NtcError) as e:
        nt.Sketch([32])
    assert e.value.code == nt.api.NTC_ENODEVICE


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ntcard_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.replace("test oracle", ""), f"{fn} mentions the oracle"


def test_create_argument_errors():
    h = ctypes.c_void_p()
    kl = (ctypes.c_uint * 1)(32)
    p = ctypes.cast(kl, ctypes.POINTER(ctypes.c_uint32))
    assert nt.lib.ntc_create(ctypes.byref(h), p, 0, 27, 7, 0, None, None) == nt.api.NTC_EINVAL
    assert nt.lib.ntc_create(ctypes.byref(h), p, 1, 0, 7, 0, None, None) == nt.api.NTC_EINVAL
    assert nt.lib.ntc_create(ctypes.byref(h), p, 1, 27, 0, 0, None, None) == nt.api.NTC_EINVAL
    assert b"rBits" in nt.lib.ntc_last_error()


def unpack(words, off):
    out = []
    for i in range(len(off) - 1):
        L = int(words[off[i]])
        w = words[off[i] + 1: off[i] + 1 + (L + 15) // 16]
        s = bytearray()
        for j in range(L):
            s.append(b"ACGT"[(int(w[j >> 4]) >> (2 * (j & 15))) & 3])
        out.append(bytes(s))
    return out


def test_pack_splits_at_invalid(oracle):
    reads = [b"ACGTNACGTACGTTTGA", b"", b"NNN", b"acguACGU" * 5, b"A", b"NACGTN", b"ACGT-ACGT.ACGTXACGT\r"]
    words, off = nt.pack_reads(reads, min_len=1)
    segs = unpack(words, off)
    assert segs == [b"ACGT", b"ACGTACGTTTGA", b"ACGTACGT" * 5, b"A", b"ACGT", b"ACGT", b"ACGT", b"ACGT", b"ACGT"]
    words, off = nt.pack_reads(reads, min_len=5)
    assert unpack(words, off) == [b"ACGTACGTTTGA", b"ACGTACGT" * 5]
    # same hashes from segments as from the raw reads (oracle), per k
    for k in (1, 4, 12):
        a = np.concatenate([oracle.hash_seq(r, k)[0] for r in reads])
        b = np.concatenate([oracle.hash_seq(s, k)[0] for s in segs] + [np.zeros(0, dtype=np.uint64)])
        assert np.array_equal(np.sort(a), np.sort(b))


def test_pack_random_bytes_against_plain_python():
    """the packer's 16-characters-at-a-time path (validity mask, 2-bit codes from the ASCII bits, partial last word) against a
    plain restatement: every byte value occurs, segment boundaries fall on every offset inside a block, record words are exact
    (unused high bits of the last word are zero)"""
    import re
    rng = np.random.default_rng(12)
    alphabet = np.frombuffer(b"ACGTUacgtu" * 40 + bytes(range(256)), dtype=np.uint8)      # ~14 % of the bytes are invalid
    clean = np.frombuffer(b"ACGTUacgtu", dtype=np.uint8)
    reads = []
    for L in list(range(0, 70)) + [150, 151, 255, 256, 257, 1000]:
        for rep in range(4):
            reads.append(bytes(rng.choice(alphabet if rep < 2 else clean, size=L)))
    tr = bytes.maketrans(b"acgtuU", b"ACGTTT")
    for min_len in (1, 12, 32):
        words, off = nt.pack_reads(reads, min_len=min_len)
        want = [seg.translate(tr) for r in reads for seg in re.split(rb"[^ACGTUacgtu]+", r) if len(seg) >= max(min_len, 1)]
        assert unpack(words, off) == want
        assert int(off[-1]) == len(words) and nt.check_offsets(off, len(words)) >= 1
        for i in range(len(off) - 1):                                  # exact words: nothing above the last base
            L = int(words[off[i]])
            assert off[i + 1] - off[i] == 1 + (L + 15) // 16
            if L % 16:
                assert int(words[off[i + 1] - 1]) >> (2 * (L % 16)) == 0


def test_pack_overflow_reports():
    chars = np.frombuffer(b"ACGTACGTACGTACGTACGT" * 3 + b"\0", dtype=np.uint8)
    soff = np.array([0, 20, 40, 60], dtype=np.uint64)
    words = np.empty(4, dtype=np.uint32)  # room for one 20-base record (1+2 words) only
    off = np.empty(8, dtype=np.uint32)
    nw, nr, cons = ctypes.c_size_t(0), ctypes.c_size_t(0), ctypes.c_size_t(0)
    rc = nt.lib.ntc_pack_seqs(chars.ctypes.data, soff.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), 3, 1,
                              words.ctypes.data, 4, ctypes.byref(nw), off.ctypes.data, 7, ctypes.byref(nr), ctypes.byref(cons))
    assert rc == nt.api.NTC_ENOMEM and cons.value == 1 and nr.value == 1 and nw.value == 3


def test_generator_matches_oracle(oracle):
    for mode, U in ((0, 0), (1, 7), (2, 0)):
        for L in (1, 31, 32, 33, 150, 400):
            a = nt.gen_ascii(3, 5, 40, L, mode, U)
            b = oracle.gen_reads(3, 5, 40, L, mode, U)
            assert np.array_equal(a, b), (mode, L)
            if mode < 2:
                stride = nt.stride_words(L)
                p = nt.gen_packed(3, 5, 40, L, mode, U, stride)
                off = np.arange(41, dtype=np.uint32) * stride
                assert unpack(p, off) == [bytes(a[i * L:(i + 1) * L]) for i in range(40)]


def test_estimate_matches_golden():
    for c in load_golden("compest_cases.json")["cases"]:
        p = np.zeros((2, 65536), dtype=np.uint32)
        for t, i, v in c["p_hist_nonzero"]:
            p[t, i] = v
        F0, f = nt.estimate(p_hist=p, rBits=c["rBits"], sBits=c["sBits"], covMax=1000)
        assert F0 == c["F0"] and [float(x) for x in f[1:]] == c["f"]
        assert f[0] == 0


def test_estimate_from_counters_matches_oracle(oracle):
    rs = np.random.RandomState(5)
    sk = (rs.poisson(0.1, size=2 << 13) * rs.randint(1, 4, size=2 << 13)).astype(np.uint16)
    F0, f = nt.estimate(t_counter=sk, rBits=13, sBits=7, covMax=200)
    oF0, of = oracle.compest(sk, None, 13, 7, 200)
    assert F0 == oF0 and np.array_equal(f[1:], of[1:201])


def test_check_offsets():
    """the offset checks of ntc_submit (host helper, no device): strictly ascending, inside the batch; longest record"""
    rng = np.random.default_rng(3)
    sizes = rng.integers(1, 20, size=100_000).astype(np.uint32)
    off = np.zeros(len(sizes) + 1, dtype=np.uint32)
    np.cumsum(sizes, out=off[1:])
    assert nt.check_offsets(off, int(off[-1])) == int(sizes.max())
    assert nt.check_offsets(off[:1], 0) == 0                       # no record
    with pytest.raises(nt.NtcError, match="off\\[n_rec\\] exceeds n_words"):
        nt.check_offsets(off, int(off[-1]) - 1)
    bad = off.copy()
    bad[70_001] = bad[70_000]                                       # record 70000 without a length word
    with pytest.raises(nt.NtcError, match="record 70000 has no length word"):
        nt.check_offsets(bad, int(off[-1]))
    bad = off.copy()
    bad[5] = bad[4] - 1                                             # descending
    with pytest.raises(nt.NtcError, match="record 4 has no length word"):
        nt.check_offsets(bad, int(off[-1]))


def test_sbits_rule():
    # ntcard.cpp:427-431
    assert nt.apply_sbits_rule(49_999_999_999) == 7 and nt.apply_sbits_rule(50_000_000_000) == 11
    assert nt.apply_sbits_rule(60_000_000_000, sBits=9) == 9


def test_bitslice_filter_selftest_on_cpu(tmp_path):
    """tools/bitslice_selftest.cu, compiled as plain C++: the bit-sliced upper-ring filter of the scan kernel
    (bitslice_core.cuh: step / sampled_mask / init_state / transpose32) marks exactly the k-mers ntComp samples
    (ntcard.cpp:132-145), for k in {5, 12, 31, 32, 63, 64, 96, 128} and several sBits, on random reads."""
    import subprocess
    exe = os.path.join(str(tmp_path), "bs_selftest")
    subprocess.run(["g++", "-O1", "-std=c++17", "-x", "c++", "-o", exe, os.path.join(ROOT, "tools", "bitslice_selftest.cu")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert "ALL OK" in out and " 0 mismatches" in out and "mismatches" in out
    assert all(" 0 mismatches" in l for l in out.splitlines() if "mismatches" in l)


def test_parallel_file_reader_equals_serial(tmp_path):
    """bin/ntcard -t N with fewer files than threads parses pieces of one FASTQ / FASTA file in parallel (reader.cpp
    read_file_parallel).  Host-only check with the stub device of tools/reader_bench: the multiset of packed records (an
    order-independent digest) must equal the serial reader's, also for a file truncated inside its last record, one without
    a trailing newline, and multi-line FASTA."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rb = os.path.join(root, "tools", "reader_bench")
    subprocess.run(["make", "-C", rb], check=True, capture_output=True)
    n, L = 90_000, 150
    a = nt.gen_ascii(7, 0, n, L, mode=2).reshape(n, L)
    rec = np.empty((n, 2 * L + 7), dtype=np.uint8)
    rec[:, 0:3] = np.frombuffer(b"@r\n", dtype=np.uint8)
    rec[:, 3:3 + L] = a
    rec[:, 3 + L:6 + L] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, 6 + L:6 + 2 * L] = np.frombuffer(b"I" * L, dtype=np.uint8)
    rec[:, 6 + 2 * L] = 10
    data = rec.tobytes()
    files = {"full.fq": data, "cut.fq": data[:-L - 1], "nonl.fq": data[:-1]}
    fa = []
    for i in range(0, n // 8):
        s = bytes(a[i * 8:(i + 1) * 8].reshape(-1))
        fa.append(b">s%d\n" % i + b"\n".join(s[j:j + 70] for j in range(0, len(s), 70)) + b"\n")
    files["multi.fa"] = b"".join(fa) * 2
    assert all(len(b) > 3 * (8 << 20) for b in files.values())  # large enough for three pieces
    for name, blob in files.items():
        p = tmp_path / name
        p.write_bytes(blob)
        outs = [subprocess.run([os.path.join(rb, "reader_bench"), "32", str(t), str(p)], check=True, capture_output=True,
                               text=True).stdout for t in (1, 2, 3)]
        assert " x 1 threads" in outs[0] and " x 2 threads" in outs[1] and " x 3 threads" in outs[2]
        tails = [o.split(");", 1)[1].split(" batches")[0].rsplit(",", 1)[0] + o.split("digest")[1] for o in outs]   # "<n> records" + digest
        assert tails[0] == tails[1] == tails[2], (name, outs)


def test_pack_bound_holds_for_the_worst_case():
    """'ANAN...' with min_len = 1 costs about one word per character -- more than the old bound
    n + total/16 + total/2 allowed, so pack_reads raised NTC_ENOMEM.  ntc_pack_bound is safe for any min_len now and
    ntc_pack_bound_k is tight enough for long sequences with a real k."""
    worst = b"AN" * 5000
    w, off = nt.pack_reads([worst], min_len=1)
    assert len(off) - 1 == 5000 and len(w) == 2 * 5000            # 5000 records of one base: length word + one base word
    assert nt.lib.ntc_pack_bound(1, len(worst)) >= len(w)
    assert nt.lib.ntc_pack_bound_k(1, len(worst), 1) >= len(w)
    for m in (2, 3, 16, 31):
        seq = (b"A" * m + b"N") * 700
        w, off = nt.pack_reads([seq], min_len=m)
        assert len(off) - 1 == 700
        assert nt.lib.ntc_pack_bound_k(1, len(seq), m) >= len(w), m
    long_seq = b"ACGT" * 250_000                                      # 1 Mbp, no N: the tight bound stays near len/16
    assert nt.lib.ntc_pack_bound_k(1, len(long_seq), 32) < len(long_seq) // 8
