"""nthll, the HyperLogLog distinct-k-mer estimator that ships beside ntcard (nthll.cpp:92-104 ntComp /
ntRead, :234-239 max merge, :243-254 the estimate).  CPU: the oracle's restatement and the product's host estimator
against the golden values made from the unmodified reference (tests/golden/hll_cases.json, oracle/make_golden.py --only
hll) and against the reference itself where oracle/_ref exists; the world-size-2 max reduction over gloo.  GPU: the
device registers (hll_kernel through the C-ABI, ntc_hll_*) bit-exact against the oracle, and bin/nthll's output line
against the unmodified reference CLI's."""
import os
import random
import subprocess
import sys

import numpy as np
import pytest

import ntcard_b200 as nt
from conftest import ROOT, load_golden

CLI = os.path.join(ROOT, "bin", "nthll")


def _reads(oracle, g):
    S, n, L, mode, U = g
    a = oracle.gen_reads(S, 0, n, L, mode, U)
    return [bytes(a[i * L:(i + 1) * L]) for i in range(n)]


def _fnv(a):
    h = 0xcbf29ce484222325
    for v in bytes(a):
        h = ((h ^ v) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return f"{h:#018x}"


def _check_case(regs, c):
    assert (_fnv(regs), int(regs.max()), int(np.count_nonzero(regs)), int(regs.astype(np.uint64).sum())) == \
        (c["digest"], c["max"], c["nonzero"], c["sum"]), c["name"]
    if "regs" in c:
        assert [int(x) for x in regs] == c["regs"]


# ---- CPU ------------------------------------------------------------------------------------------------------------
def test_oracle_hll_registers_match_reference_digests(oracle):
    for c in load_golden("hll_cases.json")["registers"]:
        regs = oracle.hll_registers(_reads(oracle, c["gen"]), c["k"], c["nBits"], nthreads=4)
        _check_case(regs, c)


def test_oracle_hll_equals_reference_on_random_input(oracle):
    from oracle.pyoracle import HllReference
    if not HllReference.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    ref = HllReference()
    rng = random.Random(99)
    reads = [bytes(rng.choice(b"ACGTacgtNu") for _ in range(rng.randint(0, 300))) for _ in range(3000)]
    for k, nBits in ((12, 8), (20, 16), (32, 16), (64, 11), (5, 2)):
        assert np.array_equal(oracle.hll_registers(reads, k, nBits, 2), ref.hll_registers(reads, k, nBits, 2)), (k, nBits)


def test_hll_estimate_matches_reference_cli_line(oracle):
    """registers by the oracle, estimate by the PRODUCT's host function -> the line the unmodified CLI printed"""
    gold = load_golden("hll_cases.json")["cli"]
    g = gold["gen_a"]
    reads = _reads(oracle, [g["S"], g["n"], g["L"], g["mode"], g["U"]])
    for tag, k, nBits, mult in (("fq_k32", 32, 16, 1), ("fq_default", 64, 16, 1), ("fq_k12_b12_t2", 12, 12, 2), ("fq_k151", 151, 16, 1)):
        regs = oracle.hll_registers(reads * mult, k, nBits, nthreads=4)
        est = nt.hll_estimate(regs, nBits)
        assert est == oracle.hll_estimate(regs, nBits)
        assert f"F0, Exp# of distnt kmers(k={k}): {int(est)}\n" == gold[tag]["stdout"], tag


def test_hll_estimate_argument_errors():
    with pytest.raises(ValueError):
        nt.hll_estimate(np.zeros(10, dtype=np.uint8), 16)
    bad = np.zeros(16, dtype=np.uint8)
    bad[3] = 64
    with pytest.raises(nt.NtcError):
        nt.hll_estimate(bad, 4)
    assert nt.lib.ntc_hll_estimate(None, 16, 1, None) == nt.api.NTC_EINVAL


def test_hll_no_cpu_fallback():
    if nt.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(nt.NtcError) as e:
        nt.HllSketch(32)
    assert e.value.code == nt.api.NTC_ENODEVICE
    r = subprocess.run([CLI, "-k32", os.devnull], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stderr and r.stdout == ""


def test_nthll_cli_usage():
    assert os.path.exists(CLI), "bin/nthll missing: run __graft_entry__.build()"
    r = subprocess.run([CLI], capture_output=True, text=True)                      # nthll.cpp:178-185
    assert r.returncode == 1 and "nthll: missing arguments" in r.stderr and "Try `nthll --help'" in r.stderr
    r = subprocess.run([CLI, "-k", "3x", "f.fq"], capture_output=True, text=True)  # nthll.cpp:173-177
    assert r.returncode == 1 and "invalid option: `-k3x'" in r.stderr
    r = subprocess.run([CLI, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "Usage: nthll [OPTION]... FILES..." in r.stderr


def _gloo_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ntcard_b200.dist import all_reduce_hll, shard_range
    from oracle.pyoracle import Oracle
    orc = Oracle()
    a = orc.gen_reads(31, 0, 4000, 120, 1, 700)
    reads = [bytes(a[i * 120:(i + 1) * 120]) for i in range(4000)]
    lo, hi = shard_range(len(reads), rank, world)
    # the checker stands in for the device: the host plumbing (shard + max all-reduce) is what is under test
    mine = torch.from_numpy(orc.hll_registers(reads[lo:hi], 32, 12))
    got = all_reduce_hll(mine).numpy()
    want = orc.hll_registers(reads, 32, 12)
    with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as f:
        f.write(f"{int(np.array_equal(got, want))}\n")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_hll_max_reduce(tmp_path):
    import torch.multiprocessing as mp
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(os.path.join(str(tmp_path), f"r{r}.txt")).read().strip() for r in range(2)] == ["1", "1"]


# ---- GPU ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_device_hll_registers_match_golden_and_oracle(oracle):
    for c in load_golden("hll_cases.json")["registers"]:
        reads = _reads(oracle, c["gen"])
        with nt.HllSketch(c["k"], c["nBits"]) as h:
            half = len(reads) // 2
            h.submit_reads(reads[:half])        # ragged batches (N-split segments, long reads cut into pieces)
            h.submit_reads(reads[half:])
            regs, n_kmers = h.finish()
            _check_case(regs, c)
            want = oracle.hll_registers(reads, c["k"], c["nBits"], nthreads=4)
            assert np.array_equal(regs, want), c["name"]
            assert n_kmers == sum(max(0, len(s) - c["k"] + 1) for r in reads
                                  for s in __import__("re").split(rb"[^ACGTUacgtu]+", r))
            assert nt.hll_estimate(regs, c["nBits"]) == oracle.hll_estimate(want, c["nBits"])
            # a second pass over the same reads changes nothing (max is idempotent); reset gives fresh registers
            h.submit_reads(reads)
            assert np.array_equal(h.finish()[0], want)
            h.reset()
            regs0, n0 = h.finish()
            assert not regs0.any() and n0 == 0


@pytest.mark.gpu
def test_device_hll_uniform_batches_and_modes(oracle):
    """uniform-stride batches (host and device resident), both register homes (shared memory <= 17 bits, global above),
    and the mode guards of the C-ABI"""
    import torch
    L, n = 150, 60000
    stride = nt.stride_words(L)
    words = nt.gen_packed(41, 0, n, L, 1, n // 6, stride)
    a = oracle.gen_reads(41, 0, n, L, 1, n // 6)
    reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
    for k, nBits in ((32, 16), (64, 17), (31, 18), (12, 20), (150, 16), (151, 16)):
        want = oracle.hll_registers(reads, k, nBits, nthreads=4)
        with nt.HllSketch(k, nBits) as h:
            h.submit(words, None, n, stride)
            regs, nk = h.finish()
            assert np.array_equal(regs, want), (k, nBits)
            assert nk == n * max(0, L - k + 1)
            h.reset()
            d = torch.from_numpy(words.view(np.int32)).cuda()
            h.submit_device(d.data_ptr(), len(words), n, stride)
            assert np.array_equal(h.finish()[0], want), (k, nBits, "device")
            with pytest.raises(nt.NtcError) as e:
                nt.Sketch.finish(h)
            assert e.value.code == nt.api.NTC_ESTATE
            with pytest.raises(nt.NtcError):
                h.counters_device()
    with nt.Sketch([32], rBits=12) as s:
        assert nt.lib.ntc_hll_finish(s.h, None, None) == nt.api.NTC_ESTATE


@pytest.mark.gpu
@pytest.mark.parametrize("k,nBits,L", [(32, 16, 150), (25, 14, 100), (64, 16, 151)])
def test_device_hll_fast_path_at_size_is_bit_exact(oracle, k, nBits, L):
    """Large uniform batches: after the first chunk every register is high enough for the bit-sliced pre-filter ("top T bits
    of the canonical hash are zero", scan kernel) + hll_hit_kernel to take over from the 64-bit recurrence.  The registers
    must stay bit-exact against the oracle (the filter has no false negatives; every candidate goes through nthll's ntComp),
    over several batches (the sharper T = 13 filter once the registers allow it), and after a reset."""
    import os
    n = 4_000_000
    stride = nt.stride_words(L)
    a = oracle.gen_reads(51, 0, n, L, 0, 0)
    off = np.arange(n + 1, dtype=np.uint64) * L
    want = np.zeros(1 << nBits, dtype=np.uint8)
    oracle.lib.orc_hll_batch(a.ctypes.data, off.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_uint64)), n, k, nBits,
                             want.ctypes.data, os.cpu_count() or 1)
    del a
    words = nt.gen_packed(51, 0, n, L, 0, 0, stride)
    with nt.HllSketch(k, nBits) as h:
        for rep in range(2):
            h.reset()
            half = n // 2
            h.submit(words[:half * stride], None, half, stride)          # chunks through the recurrence, then the filter
            h.submit(words[half * stride:], None, n - half, stride)      # filter from the start
            regs, nk = h.finish()
            assert nk == n * (L - k + 1)
            assert np.array_equal(regs, want), (k, nBits, rep, int((regs != want).sum()))
        launches = h.stats()["launches"]
    assert launches >= 8  # (general chunk + min) + (scan + hit) ... per pass: the fast path did run


@pytest.mark.gpu
@pytest.mark.parametrize("k,nBits", [(32, 12), (40, 16)])
def test_device_hll_fast_path_with_records_of_mixed_lengths(oracle, k, nBits):
    """Uniform-stride batches whose records differ in length (some shorter than k): the pre-filter scans such tiles at the longest
    length and hll_hit_kernel drops the candidates that lie past a record's own end.  nBits = 12: the registers pass 12 within the
    first batch, so the second one takes the fixed 13-bit filter; nBits = 16: levels picked on the device throughout."""
    import ctypes
    import os
    n, L = 800_000, 150
    stride = nt.stride_words(L)
    a = oracle.gen_reads(77, 0, n, L, 0, 0).reshape(n, L)
    rng = np.random.default_rng(5)
    lens = np.where(rng.random(n) < 0.3, rng.integers(20, L + 1, n), L).astype(np.uint64)
    keep = np.arange(L, dtype=np.uint64)[None, :] < lens[:, None]
    buf = np.ascontiguousarray(a[keep])
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    want = np.zeros(1 << nBits, dtype=np.uint8)
    oracle.lib.orc_hll_batch(buf.ctypes.data, off.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), n, k, nBits, want.ctypes.data, os.cpu_count() or 1)
    w, woff = nt.pack_chars(buf, off, min_len=1)
    assert len(woff) - 1 == n
    cnt = np.diff(woff.astype(np.int64))
    rows = np.zeros((n, stride), dtype=np.uint32)
    rows[np.arange(stride)[None, :] < cnt[:, None]] = w      # record i = its length word + packed bases, zero padding behind
    rows = rows.reshape(-1)
    with nt.HllSketch(k, nBits) as h:
        half = n // 2
        h.submit(rows[:half * stride], None, half, stride)
        l0 = h.stats()["launches"]
        h.submit(rows[half * stride:], None, n - half, stride)
        assert h.stats()["launches"] - l0 <= 3 * 4                  # scan + hit (+ min) per chunk: no recurrence in the second batch
        regs, nk = h.finish()
    assert nk == int(np.maximum(lens.astype(np.int64) - k + 1, 0).sum())
    assert np.array_equal(regs, want), (k, nBits, int((regs != want).sum()))


@pytest.mark.gpu
def test_device_hll_caller_owned_registers_cleared_by_the_caller(oracle):
    """Registers in a torch tensor that the caller clears between two batches WITHOUT ntc_reset: the pre-filter must not go on
    assuming the high registers of the first batch (it re-reads their minimum once per batch when it does not own them)."""
    import ctypes
    import os
    import torch
    n, L, k, nBits = 400_000, 150, 32, 10
    stride = nt.stride_words(L)
    t = torch.zeros(1 << nBits, dtype=torch.uint8, device="cuda")
    words = nt.gen_packed(91, 0, 2 * n, L, 0, 0, stride)
    a = oracle.gen_reads(91, n, n, L, 0, 0)
    off = np.arange(n + 1, dtype=np.uint64) * L
    want = np.zeros(1 << nBits, dtype=np.uint8)
    oracle.lib.orc_hll_batch(a.ctypes.data, off.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), n, k, nBits, want.ctypes.data, os.cpu_count() or 1)
    with nt.HllSketch(k, nBits, d_regs=t.data_ptr()) as h:
        h.submit(words[:n * stride], None, n, stride)
        h.sync()
        assert int(t.min()) >= 12                                  # the fixed 13-bit filter would be next
        t.zero_()
        torch.cuda.synchronize()
        h.submit(words[n * stride:], None, n, stride)
        h.sync()
        assert np.array_equal(t.cpu().numpy(), want)


@pytest.mark.gpu
def test_device_hll_caller_owned_registers_merge_by_max(oracle):
    """two contexts on one device, sharded reads, registers in torch tensors, merged by max = one context over all reads
    (the N>1 path without a second GPU: ntcard_b200.dist.all_reduce_hll does torch.maximum's job across ranks)"""
    import torch
    a = oracle.gen_reads(43, 0, 20000, 150, 0, 0)
    reads = [bytes(a[i * 150:(i + 1) * 150]) for i in range(20000)]
    t = [torch.zeros(1 << 16, dtype=torch.uint8, device="cuda") for _ in range(2)]
    hs = [nt.HllSketch(32, 16, d_regs=x.data_ptr()) for x in t]
    hs[0].submit_reads(reads[:10000])
    hs[1].submit_reads(reads[10000:])
    for h in hs:
        h.sync()
        assert h.registers_device() == (t[hs.index(h)].data_ptr(), 1 << 16)
    merged = torch.maximum(t[0], t[1]).cpu().numpy()
    assert np.array_equal(merged, oracle.hll_registers(reads, 32, 16, nthreads=4))
    for h in hs:
        h.close()


@pytest.mark.gpu
def test_nthll_cli_matches_reference_cli(tmp_path, oracle):
    gold = load_golden("hll_cases.json")["cli"]
    g = gold["gen_a"]
    reads = [r.decode() for r in _reads(oracle, [g["S"], g["n"], g["L"], g["mode"], g["U"]])]
    td = str(tmp_path)
    with open(os.path.join(td, "a.fq"), "w") as f:
        for i, r in enumerate(reads):
            f.write(f"@r{i}\n{r}\n+\n{'I' * len(r)}\n")
    with open(os.path.join(td, "a.fa"), "w") as f:
        for i, r in enumerate(reads):
            f.write(f">r{i}\n" + "\n".join(r[j:j + 60] for j in range(0, len(r), 60)) + "\n")
    with open(os.path.join(td, "a.sam"), "w") as f:
        f.write("@HD\tVN:1.6\n@SQ\tSN:x\tLN:1000\n")
        for i, r in enumerate(reads):
            f.write(f"r{i}\t4\t*\t0\t0\t*\t*\t0\t0\t{r}\t{'I' * len(r)}\n")
    for tag, c in gold.items():
        if tag == "gen_a":
            continue
        r = subprocess.run([CLI] + c["args"] + [os.path.join(td, x) for x in c["files"]], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout == c["stdout"], tag
    # @list argument and a missing file (silently skipped, nthll.cpp:219-229)
    with open(os.path.join(td, "list.txt"), "w") as f:
        f.write(os.path.join(td, "a.fq") + "\n" + os.path.join(td, "nope.fq") + "\n")
    r = subprocess.run([CLI, "-k32", "@" + os.path.join(td, "list.txt")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == gold["fq_k32"]["stdout"]
