"""Parity tests proper: the CUDA path, called through the C-ABI, against the oracle and the golden
fixtures on the same inputs.  Integer results (sketch counters, F1) must be bit-exact; F0 / f_i are
computed by the product's own estimator and compared for equality with the oracle's compEst on the
same sketch (tolerance 0 -- tighter than the 1e-6 relative the north star allows)."""
import random

import numpy as np
import pytest

import ntcard_b200 as nt
from conftest import load_golden

pytestmark = pytest.mark.gpu

KERNELS = [nt.KERNEL_ROLL64, nt.KERNEL_AUTO]


def H(x):
    return int(x, 16)


def reads_of(case, oracle):
    if "reads" in case:
        return [r.encode("latin1") for r in case["reads"]]
    g = case["gen"]
    a = oracle.gen_reads(g["S"], 0, g["n"], g["L"], g["mode"], g["U"])
    reads = [bytes(a[i * g["L"]:(i + 1) * g["L"]]) for i in range(g["n"])]
    return reads * g.get("repeat", 1)


def run_gpu(reads, kList, rBits, sBits, kernel, batches=1):
    with nt.Sketch(kList, rBits=rBits, sBits=sBits) as sk:
        sk.set_kernel(kernel)
        step = (len(reads) + batches - 1) // batches
        for b in range(0, len(reads), max(step, 1)):
            sk.submit_reads(reads[b:b + step])
        t, f1, p = sk.finish(counters=True, hist=True)
    return t, f1, p


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", ["FX1", "small_r12_s3", "nmode_s7", "nmode_s11", "wrap_u16", "ragged_r16_s2"])
def test_golden_sketch_cases(oracle, name, kernel):
    case = next(c for c in load_golden("sketch_cases.json")["cases"] if c["name"] == name)
    reads = reads_of(case, oracle)
    rB = 1 << case["rBits"]
    t, f1, p = run_gpu(reads, case["k"], case["rBits"], case["sBits"], kernel, batches=3)
    assert [int(x) for x in f1] == case["F1"]
    flat = t.reshape(-1, rB)
    for i, gt in enumerate(case["tables"]):
        tab = np.ascontiguousarray(flat[i])
        d = oracle.table_digest(tab)
        assert (d["nnz"], d["sum"], d["max"]) == (gt["nnz"], gt["sum"], gt["max"])
        assert d["digest"] == H(gt["digest"])
        assert np.array_equal(p.reshape(-1, 65536)[i], np.bincount(tab, minlength=65536))
    for ki, ge in enumerate(case["est"]):
        F0, f = nt.estimate(p_hist=p[ki], rBits=case["rBits"], sBits=case["sBits"], covMax=64)
        assert F0 == ge["F0"] and [float(x) for x in f[1:]] == ge["f"]


@pytest.mark.parametrize("kernel", KERNELS)
def test_random_ragged_vs_oracle(oracle, kernel):
    rng = random.Random(77)
    reads = []
    for _ in range(4000):
        L = rng.choice((0, 1, 11, 12, 31, 32, 33, 64, 65, 100, 150, 151, 300, 700, 2500))
        reads.append(bytes(rng.choice(b"ACGTNacgtuX") if rng.random() < 0.02 else rng.choice(b"ACGT") for _ in range(L)))
    for kList, rBits, sBits in (([12, 32, 33, 64, 128], 16, 3), ([31], 14, 1), ([20, 96], 18, 7)):
        want, wf1 = oracle.sketch_reads(reads, kList, rBits, sBits, nthreads=4)
        t, f1, _ = run_gpu(reads, kList, rBits, sBits, kernel, batches=2)
        assert np.array_equal(f1, wf1)
        assert np.array_equal(t.reshape(-1), want)


@pytest.mark.parametrize("kernel", KERNELS)
def test_long_reads_with_n_vs_oracle(oracle, kernel):
    # config 5 in miniature: 10 kbp reads with N runs, k=31, s=11 -> pieces with k-1 overlap
    a = oracle.gen_reads(4, 0, 300, 10000, 2, 0)
    reads = [bytes(a[i * 10000:(i + 1) * 10000]) for i in range(300)]
    for sBits in (11, 4):
        want, wf1 = oracle.sketch_reads(reads, [31, 250], 20, sBits, nthreads=4)
        t, f1, _ = run_gpu(reads, [31, 250], 20, sBits, kernel)
        assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want)


@pytest.mark.parametrize("kernel", KERNELS)
def test_uniform_stride_batches(oracle, kernel):
    # the packed uniform layout used by the bench (stride multiple of 4 words and not)
    for L, mode, U in ((150, 0, 0), (150, 1, 500), (100, 1, 100), (64, 0, 0), (33, 0, 0)):
        n = 6000
        a = oracle.gen_reads(9, 0, n, L, mode, U)
        reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
        kList = [k for k in (12, 32, 64) if k <= L] or [12]
        want, wf1 = oracle.sketch_reads(reads, kList, 18, 5, nthreads=4)
        for align4 in (True, False):
            stride = nt.stride_words(L, align4)
            words = nt.gen_packed(9, 0, n, L, mode, U, stride)
            with nt.Sketch(kList, rBits=18, sBits=5) as sk:
                sk.set_kernel(kernel)
                sk.submit(words[: (n // 2) * stride], None, n // 2, stride)
                sk.submit(words[(n // 2) * stride:], None, n - n // 2, stride)
                t, f1, _ = sk.finish(counters=True, hist=False)
            assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want), (L, mode, align4)


def test_device_generator_and_device_submit(oracle):
    import torch
    L, n, U = 150, 50000, 5000
    stride = nt.stride_words(L)
    a = oracle.gen_reads(12, 100, n, L, 1, U)
    reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
    want, wf1 = oracle.sketch_reads(reads, [32], 20, 7, nthreads=4)
    d = torch.empty(n * stride, dtype=torch.int32, device="cuda")
    with nt.Sketch([32], rBits=20, sBits=7) as sk:
        sk.gen_packed_device(12, 100, n, L, 1, U, stride, d.data_ptr())
        sk.submit_device(d.data_ptr(), n * stride, n, stride)
        t, f1, _ = sk.finish(counters=True, hist=False)
        host = nt.gen_packed(12, 100, n, L, 1, U, stride)
        assert np.array_equal(d.cpu().numpy().view(np.uint32), host)
    assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want)


def test_pinned_submit_external_counters_and_stream(oracle):
    import torch
    L, n = 150, 20000
    stride = nt.stride_words(L)
    a = oracle.gen_reads(2, 0, n, L, 0, 0)
    reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
    want, wf1 = oracle.sketch_reads(reads, [32, 64], 20, 7, nthreads=4)
    buf = nt.PinnedBuffer(n * stride)
    nt.gen_packed(2, 0, n, L, 0, 0, stride, out=buf.array)
    ctr = torch.full((2 * 2 << 20,), 7, dtype=torch.int32, device="cuda")   # garbage: ntc_create must zero it
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with nt.Sketch([32, 64], rBits=20, sBits=7, d_counters=ctr.data_ptr(), stream=st.cuda_stream) as sk:
            tk = sk.submit(buf.array, None, n, stride)
            sk.wait(tk)
            sk.sync()
            p, cnt = sk.counters_device()
            assert p == ctr.data_ptr() and cnt == ctr.numel()
            st.synchronize()
            dev = ctr.cpu().numpy().view(np.uint32)
            assert np.array_equal((dev & 0xFFFF).astype(np.uint16), want)
            t, f1, _ = sk.finish(counters=True, hist=False)
    assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want)
    buf.free()


def test_empty_and_short_inputs():
    with nt.Sketch([32], rBits=12, sBits=7) as sk:
        sk.submit_reads([])
        sk.submit_reads([b"", b"ACGT", b"N" * 100, b"ACGTACGTAC" * 3 + b"A"])   # all shorter than k after the split
        t, f1, p = sk.finish(counters=True, hist=True)
        assert int(f1[0]) == 0 and not t.any() and p[0, 0, 0] == 4096 and p[0, 1, 0] == 4096
        F0, f = nt.estimate(p_hist=p[0], rBits=12, sBits=7, covMax=10)
        assert F0 == 0 and not f.any()      # compEst early-out, ntcard.cpp:262-264


def test_reset_and_reuse(oracle):
    a = oracle.gen_reads(3, 0, 3000, 100, 0, 0)
    reads = [bytes(a[i * 100:(i + 1) * 100]) for i in range(3000)]
    want, wf1 = oracle.sketch_reads(reads, [20], 14, 2)
    with nt.Sketch([20], rBits=14, sBits=2) as sk:
        for _ in range(3):
            sk.reset()
            sk.submit_reads(reads)
            t, f1, _ = sk.finish(counters=True, hist=False)
            assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want)


@pytest.mark.parametrize("kernel", KERNELS)
def test_full_size_properties_config2(kernel):
    """BASELINE config 2 at full size (10 M x 150 bp, k=32, s=7, r=27), device-resident input.
    Size-independent properties: F1 is analytic; the sketch is additive over batches; the canonical
    hash makes a read and its reverse complement indistinguishable, so a repeat-mode input whose
    second half is the reverse complement of the first gives exactly twice the half sketch."""
    import torch
    L, n, k, stride = 150, 10_000_000, 32, nt.stride_words(150)
    d = torch.empty(n * stride, dtype=torch.int32, device="cuda")
    with nt.Sketch([k], rBits=27, sBits=7) as sk:
        sk.set_kernel(kernel)
        sk.gen_packed_device(1, 0, n, L, 1, n // 2, stride, d.data_ptr())     # read i+n/2 = revcomp(read i)
        sk.submit_device(d.data_ptr(), n * stride, n, stride)
        f1 = sk.totals()
        assert int(f1[0]) == n * (L - k + 1)
        t_full, _, p_full = sk.finish(counters=True, hist=True)
        sk.reset()
        half = n // 2
        sk.submit_device(d.data_ptr(), half * stride, half, stride)
        t_half, f1h, _ = sk.finish(counters=True, hist=False)
        assert int(f1h[0]) == half * (L - k + 1)
        assert np.array_equal(t_full, (t_half.astype(np.uint32) * 2).astype(np.uint16))
        # estimator sanity on the full sketch: every distinct k-mer occurs exactly twice
        F0, f = nt.estimate(p_hist=p_full[0], rBits=27, sBits=7, covMax=4)
        distinct = half * (L - k + 1)
        assert abs(F0 - distinct) / distinct < 0.02 and f[2] > 0.97 * distinct and f[1] < 0.01 * distinct


# ---- the bit-sliced kernel (NTC_KERNEL_BITSLICE), forced --------------------------------------
@pytest.mark.parametrize("sBits", [7, 11])
@pytest.mark.parametrize("L,kList", [(150, [32]), (150, [12, 31, 64, 96, 128]), (100, [32, 64]), (151, [31, 32]),
                                     (64, [32, 64]), (40, [12, 32]), (170, [32])])
def test_bitslice_uniform_reads(oracle, L, kList, sBits):
    n = 5000 if sBits == 7 else 40000      # partial last tile (n % 1024 != 0) takes the in-kernel general path
    a = oracle.gen_reads(31, 0, n, L, 1, n // 4)
    reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
    want, wf1 = oracle.sketch_reads(reads, kList, 20, sBits, nthreads=4)
    stride = nt.stride_words(L)
    words = nt.gen_packed(31, 0, n, L, 1, n // 4, stride)
    with nt.Sketch(kList, rBits=20, sBits=sBits) as sk:
        sk.set_kernel(nt.KERNEL_BITSLICE)
        sk.submit(words, None, n, stride)
        t, f1, _ = sk.finish(counters=True, hist=False)
    assert np.array_equal(f1, wf1)
    assert np.array_equal(t.reshape(-1), want)


@pytest.mark.parametrize("L,sBits,n,kList", [(150, 7, 5000, list(range(13, 29))), (150, 7, 5000, list(range(29, 44))),
                                            (150, 11, 40000, [21, 25, 41, 51, 55, 71, 77, 91, 101, 121, 127, 150]),
                                            (250, 7, 5000, [21, 33, 45, 57, 69, 81, 111, 151, 171, 200, 222, 250]),
                                            (300, 11, 40000, [35, 99, 160, 201, 255, 287, 288, 300])])
def test_bitslice_every_k_class(oracle, L, sBits, n, kList):
    """every k mod 31 variant of the scan kernel (Makefile BS_KMS = 0..30) against the oracle, forced: k = 13..43 covers the
    31 classes once, the other lists are the usual assembly k values and k up to the record length (k >= 288: general kernel)"""
    a = oracle.gen_reads(37, 0, n, L, 1, n // 4)
    reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
    want, wf1 = oracle.sketch_reads(reads, kList, 18, sBits, nthreads=4)
    stride = nt.stride_words(L)
    words = nt.gen_packed(37, 0, n, L, 1, n // 4, stride)
    with nt.Sketch(kList, rBits=18, sBits=sBits) as sk:
        if max(kList) < 288:
            sk.set_kernel(nt.KERNEL_BITSLICE)       # fails loudly if a class has no variant
        sk.submit(words, None, n, stride)
        t, f1, _ = sk.finish(counters=True, hist=False)
    assert np.array_equal(f1, wf1)
    assert np.array_equal(t.reshape(-1), want)


def test_bitslice_mixed_lengths_in_uniform_stride(oracle):
    """Records of different lengths inside a uniform-stride batch: tiles that are not uniform fall back
    to the 64-bit path inside the kernel; results stay exact."""
    rng = random.Random(5)
    n, stride = 6000, 12
    reads = []
    for i in range(n):
        L = 150 if (i // 1024) % 2 == 0 else rng.choice((20, 31, 32, 100, 150, 176))
        reads.append(bytes(rng.choice(b"ACGT") for _ in range(L)))
    words = np.zeros(n * stride, dtype=np.uint32)
    for i, r in enumerate(reads):
        w, _ = nt.pack_reads([r])
        words[i * stride:i * stride + len(w)] = w
        if len(r) == 0:
            words[i * stride] = 0
    for kList in ([32], [31, 64]):
        want, wf1 = oracle.sketch_reads(reads, kList, 18, 7, nthreads=4)
        with nt.Sketch(kList, rBits=18, sBits=7) as sk:
            sk.set_kernel(nt.KERNEL_BITSLICE)
            sk.submit(words, None, n, stride)
            t, f1, _ = sk.finish(counters=True, hist=False)
        assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want)


@pytest.mark.parametrize("sBits,kList", [(7, [32, 44, 64, 73]), (11, [21, 31, 55, 128]), (7, [12, 96, 150])])
def test_bitslice_ragged_lengths_in_every_tile(oracle, sBits, kList):
    """Trimmed reads: every record of every tile has its own length (same stride).  The scan kernel runs such tiles at the
    tile's longest length and the hit kernel drops what lies past a record's end; for the k whose all-A k-mer is itself
    sampled (k = 44, 73 at sBits 7: the zero padding would flood the hit queue) the tile goes to the fallback kernel.
    Padding beyond a record's last base is zero in one half of the batch and random in the other."""
    rng = random.Random(11)
    n, stride = 7000, 12
    lens = [rng.choice((0, 5, 31, 32, 33, 63, 64)) if rng.random() < 0.05 else rng.randint(90, 176) for _ in range(n)]
    reads = [bytes(rng.choice(b"ACGT") for _ in range(L)) for L in lens]
    words = np.zeros(n * stride, dtype=np.uint32)
    for i, r in enumerate(reads):
        w, _ = nt.pack_reads([r])
        if i >= n // 2:   # garbage behind the record: whole words, and the unused bits of its last word
            words[i * stride + 1:(i + 1) * stride] = np.frombuffer(rng.randbytes(4 * (stride - 1)), dtype=np.uint32)
            if len(r) % 16 and len(w) > 1:
                w = w.copy()
                w[-1] |= np.uint32((rng.getrandbits(32) << (2 * (len(r) % 16))) & 0xFFFFFFFF)
        words[i * stride:i * stride + len(w)] = w
        words[i * stride] = len(r)
    want, wf1 = oracle.sketch_reads(reads, kList, 18, sBits, nthreads=4)
    with nt.Sketch(kList, rBits=18, sBits=sBits) as sk:
        sk.set_kernel(nt.KERNEL_BITSLICE)
        sk.submit(words, None, n, stride)
        t, f1, _ = sk.finish(counters=True, hist=False)
    assert np.array_equal(f1, wf1)
    assert np.array_equal(t.reshape(-1), want)


@pytest.mark.parametrize("sBits,kList", [(7, [12, 32, 44, 64]), (11, [31, 96])])
def test_ragged_short_records_are_padded_for_the_pipeline(oracle, sBits, kList):
    """The documented drop-in path (ntc_pack_seqs -> ntc_submit with offsets): reads with N runs and reads of odd lengths make a
    RAGGED batch; the library pads it to one stride on the device and the pipeline takes it as mixed-length tiles.  Forced
    NTC_KERNEL_BITSLICE proves the pipeline took every k (it fails loudly otherwise)."""
    rng = random.Random(23)
    a = oracle.gen_reads(5, 0, 6000, 150, 2, 0)                       # N mode: reads split into segments
    reads = [bytes(a[i * 150:(i + 1) * 150]) for i in range(6000)]
    reads += [bytes(rng.choice(b"ACGTacgtu") for _ in range(rng.randint(40, 240))) for _ in range(3000)]   # odd lengths, no N
    rng.shuffle(reads)
    want, wf1 = oracle.sketch_reads(reads, kList, 18, sBits, nthreads=4)
    with nt.Sketch(kList, rBits=18, sBits=sBits) as sk:
        sk.set_kernel(nt.KERNEL_BITSLICE)
        sk.submit_reads(reads[:4500])
        sk.submit_reads(reads[4500:])
        t, f1, _ = sk.finish(counters=True, hist=False)
    assert np.array_equal(f1, wf1)
    assert np.array_equal(t.reshape(-1), want)


def test_bitslice_low_complexity_queue_overflow(oracle):
    """Poly-A / dinucleotide reads: every k-mer of a read has the same hash, so either none or ALL of
    them are sampled -- the per-body hit queue overflows and the slow exact path must take over."""
    n, L, stride = 4096, 150, 12
    pats = [b"A", b"C", b"AC", b"AG", b"ACG", b"AAT", b"ACGT", b"T"]
    reads = [(pats[i % len(pats)] * 200)[:L] for i in range(n)]
    words = np.zeros(n * stride, dtype=np.uint32)
    for i, r in enumerate(reads):
        w, _ = nt.pack_reads([r])
        words[i * stride:i * stride + len(w)] = w
    for sBits in (7, 11):
        want, wf1 = oracle.sketch_reads(reads, [12, 32], 16, sBits, nthreads=4)
        with nt.Sketch([12, 32], rBits=16, sBits=sBits) as sk:
            sk.set_kernel(nt.KERNEL_BITSLICE)
            sk.submit(words, None, n, stride)
            t, f1, _ = sk.finish(counters=True, hist=False)
        assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want)


def test_bitslice_refuses_unsupported():
    with nt.Sketch([32], rBits=12, sBits=5) as sk:      # sBits 5 has no bit-sliced variant
        sk.set_kernel(nt.KERNEL_BITSLICE)
        words = nt.gen_packed(1, 0, 2048, 150, 0, 0, 12)
        with pytest.raises(nt.NtcError):
            sk.submit(words, None, 2048, 12)


def test_hist_range_matches_bincount(oracle):
    """ntc_hist_range (the per-rank histogram of a reduce-scattered slice) against numpy on the same counters."""
    import torch
    a = oracle.gen_reads(8, 0, 20000, 150, 1, 4000)
    reads = [bytes(a[i * 150:(i + 1) * 150]) for i in range(20000)]
    rBits = 18
    with nt.Sketch([20, 32], rBits=rBits, sBits=3) as sk:
        sk.submit_reads(reads)
        sk.sync()
        ptr, n = sk.counters_device()
        t, _, p_full = sk.finish(counters=True, hist=True)
        parts = np.zeros_like(p_full)
        world = 4
        for r in range(world):
            parts += sk.hist_range(ptr + 4 * r * (n // world), r * (n // world), n // world)
        parts[:, :, 0] = (1 << rBits) - parts[:, :, 1:].sum(axis=2)
        assert np.array_equal(parts, p_full)
        for ki in range(2):
            for tb in range(2):
                assert np.array_equal(p_full[ki, tb], np.bincount(t[ki, tb], minlength=65536))
        with pytest.raises(nt.NtcError):
            sk.hist_range(ptr, 100, 1000)   # not chunk aligned


# ---- the scan -> hit log -> apply pipeline: flush modes, pool exhaustion, mixed kernels ------------
def _uniform_case(oracle, seed, n, L, kList, rBits, sBits, mode=1, U=None):
    U = U or max(1, n // 4)
    a = oracle.gen_reads(seed, 0, n, L, mode, U)
    reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
    want, wf1 = oracle.sketch_reads(reads, kList, rBits, sBits, nthreads=4)
    stride = nt.stride_words(L)
    words = nt.gen_packed(seed, 0, n, L, mode, U, stride)
    return words, stride, want, wf1


@pytest.mark.parametrize("rBits", [20, 24])
def test_pipeline_many_batches_and_flushes(oracle, rBits):
    """Batches interleaved with explicit flushes: the first flush writes the zeros (state 0), the later ones
    accumulate into the materialised sketch (L2-prefetch mode); counters read through ntc_counters_device in
    between must be complete.  rBits = 24 gives several slices per k."""
    import torch
    n, L, kList = 8192, 150, [32, 64]
    words, stride, want, wf1 = _uniform_case(oracle, 41, n, L, kList, rBits, 7)
    with nt.Sketch(kList, rBits=rBits, sBits=7) as sk:
        sk.set_kernel(nt.KERNEL_BITSLICE)
        per = 2048
        for b in range(0, n, per):
            sk.submit(words[b * stride:(b + per) * stride], None, per, stride)
            if b == 0:
                sk.flush()
            if b == 2 * per:
                sk.sync()
                sk.counters_device()
        t, f1, _ = sk.finish(counters=True, hist=False)
        assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want)
        # a second pass into the same (materialised) sketch doubles every counter
        for b in range(0, n, per):
            sk.submit(words[b * stride:(b + per) * stride], None, per, stride)
        t2, f2, _ = sk.finish(counters=True, hist=False)
        assert np.array_equal(f2, 2 * wf1)
        assert np.array_equal(t2.reshape(-1), (want.astype(np.uint32) * 2).astype(np.uint16))


def test_pipeline_pool_exhaustion_is_exact(oracle, monkeypatch):
    """A hit log far too small for the batch (64 blocks): the conditional flush in front of the hit kernel must
    materialise the sketch and the hits that find no block must be added directly -- same counters."""
    monkeypatch.setenv("NTC_POOL_BLOCKS", "64")
    n, L, kList = 20000, 150, [32]
    words, stride, want, wf1 = _uniform_case(oracle, 43, n, L, kList, 22, 7)
    with nt.Sketch(kList, rBits=22, sBits=7) as sk:
        sk.set_kernel(nt.KERNEL_BITSLICE)
        sk.submit(words[:8192 * stride], None, 8192, stride)
        sk.submit(words[8192 * stride:], None, n - 8192, stride)
        t, f1, _ = sk.finish(counters=True, hist=False)
    assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want)


def test_pipeline_mixed_with_roll64_batches(oracle):
    """Uniform batches (pipeline) and ragged batches (roll64, direct increments) into one sketch, in both
    orders, with a reset in between: the lazily zeroed sketch must be materialised before the first direct
    increment and nothing may be lost or doubled."""
    n, L, kList = 4096, 150, [31, 32]
    words, stride, want_u, wf1_u = _uniform_case(oracle, 47, n, L, kList, 21, 7)
    a = oracle.gen_reads(48, 0, 3000, 120, 2, 0)
    ragged = [bytes(a[i * 120:(i + 1) * 120]) for i in range(3000)]
    want_r, wf1_r = oracle.sketch_reads(ragged, kList, 21, 7, nthreads=4)
    want = (want_u.astype(np.uint32) + want_r.astype(np.uint32)).astype(np.uint16)
    with nt.Sketch(kList, rBits=21, sBits=7) as sk:
        for order in (0, 1):
            sk.reset()
            if order == 0:
                sk.submit(words, None, n, stride)
                sk.submit_reads(ragged)
            else:
                sk.submit_reads(ragged)
                sk.submit(words, None, n, stride)
            t, f1, _ = sk.finish(counters=True, hist=False)
            assert np.array_equal(f1, wf1_u + wf1_r) and np.array_equal(t.reshape(-1), want)


def test_pipeline_general_hit_kernel_long_reads(oracle):
    """Reads too long for the staged (shared-memory) hit kernel and multi-tile units at s = 11."""
    for L, sBits, n in ((400, 7, 3072), (300, 11, 5000)):
        words, stride, want, wf1 = _uniform_case(oracle, 53, n, L, [32, 96], 20, sBits)
        with nt.Sketch([32, 96], rBits=20, sBits=sBits) as sk:
            sk.set_kernel(nt.KERNEL_BITSLICE)
            sk.submit(words, None, n, stride)
            t, f1, _ = sk.finish(counters=True, hist=False)
        assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want)


def test_long_ragged_records_are_retiled_exactly(oracle):
    """Long ragged records (long reads with N runs, mixed with short ones) are cut on the device into equal pieces for
    the pipeline plus ragged tails for the general kernel; every k-mer must still be counted exactly once, for every k."""
    rng = random.Random(123)
    a = oracle.gen_reads(21, 0, 400, 6000, 2, 0)
    reads = [bytes(a[i * 6000:(i + 1) * 6000]) for i in range(400)]
    for _ in range(300):                                        # odd lengths around the piece geometry
        L = rng.choice((0, 5, 31, 32, 145, 146, 147, 175, 176, 177, 321, 322, 323, 700, 1500, 2999))
        reads.append(bytes(rng.choice(b"ACGT") for _ in range(L)))
    rng.shuffle(reads)
    for kList, rBits, sBits in (([31], 20, 11), ([12, 32, 64], 18, 7), ([32, 96, 128], 18, 7), ([31, 33], 16, 7)):
        want, wf1 = oracle.sketch_reads(reads, kList, rBits, sBits, nthreads=4)
        t, f1, _ = run_gpu(reads, kList, rBits, sBits, nt.KERNEL_AUTO, batches=2)
        assert np.array_equal(f1, wf1), (kList, f1, wf1)
        assert np.array_equal(t.reshape(-1), want), kList


@pytest.mark.parametrize("n,kList,sBits,name", [(100_000_000, [32, 64, 96, 128], 7, "config 3"),
                                                (125_000_000, [64], 11, "config 4, one GPU's shard")])
def test_full_size_configs_pipeline_equals_general_kernel(n, kList, sBits, name):
    """BASELINE configs 3 and 4 at full size (device-resident synthetic reads, r=27).  Size-independent properties: F1 is
    analytic for every k; the two independent device implementations -- the scan/hit/apply pipeline and the 64-bit
    general kernel, each checked against the oracle at small sizes -- must give the same counter-value histogram
    (and therefore the same F0 / f_i); the histogram must account for every bucket."""
    import torch
    L, stride = 150, nt.stride_words(150)
    d = torch.empty(n * stride, dtype=torch.int32, device="cuda")
    hists = {}
    with nt.Sketch(kList, rBits=27, sBits=sBits) as sk:
        sk.gen_packed_device(2, 0, n, L, 0, 0, stride, d.data_ptr())
        for kernel in (nt.KERNEL_AUTO, nt.KERNEL_ROLL64):
            sk.reset()
            sk.set_kernel(kernel)
            half = (n // 2 // 1024) * 1024 + 7                      # two batches, the first ends inside a tile
            sk.submit_device(d.data_ptr(), half * stride, half, stride)
            sk.submit_device(d.data_ptr() + half * stride * 4, (n - half) * stride, n - half, stride)
            _, f1, p = sk.finish(counters=False, hist=True)
            assert [int(x) for x in f1] == [n * (L - k + 1) for k in kList], name
            hists[kernel] = p
    a, b = hists[nt.KERNEL_AUTO], hists[nt.KERNEL_ROLL64]
    assert np.array_equal(a, b), name
    assert (a.sum(axis=2, dtype=np.uint64) == (1 << 27)).all()
    for ki, k in enumerate(kList):
        F0, f = nt.estimate(p_hist=a[ki], rBits=27, sBits=sBits, covMax=4)
        distinct = n * (L - k + 1)                                   # uniform random reads: (almost) every k-mer is new
        assert abs(F0 - distinct) / distinct < 0.02, (name, k, F0, distinct)


def test_hit_log_exchange_between_two_contexts_on_one_gpu(oracle):
    """The multi-GPU sparse reduction (ntc_log_counts / export / import / flush_slices / hist_slices) exercised on ONE
    device: two contexts play two ranks, each sketches half of the reads, they swap the log blocks of the slices the
    other one owns (slice s belongs to rank s % 2) through plain device buffers, each materialises and histograms only
    its own slices; the summed histogram must equal the histogram of one sketch of all reads."""
    import torch
    from ntcard_b200.dist import exchange_feasible, plan_exchange
    n, L, kList, rBits, sBits = 40960, 150, [32, 64], 25, 7
    stride = nt.stride_words(L)
    words = nt.gen_packed(61, 0, n, L, 1, n // 4, stride)
    a = oracle.gen_reads(61, 0, n, L, 1, n // 4)
    reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
    want, wf1 = oracle.sketch_reads(reads, kList, rBits, sBits, nthreads=4)
    rB = 1 << rBits
    want_p = np.stack([np.bincount(want[t * rB:(t + 1) * rB], minlength=65536) for t in range(2 * len(kList))]).reshape(len(kList), 2, 65536)
    half = n // 2
    W = nt.Sketch.WIRE_BLOCK_WORDS
    with nt.Sketch(kList, rBits=rBits, sBits=sBits) as s0, nt.Sketch(kList, rBits=rBits, sBits=sBits) as s1:
        ranks = [s0, s1]
        s0.submit(words[:half * stride], None, half, stride)
        s1.submit(words[half * stride:], None, n - half, stride)
        got = [r.log_counts() for r in ranks]
        assert all(ok for _, ok, _ in got)
        counts = np.stack([c.astype(np.int64) for c, _, _ in got])
        infos = np.array([i for _, _, i in got], dtype=np.int64)
        n_slices = counts.shape[1]
        assert n_slices == len(kList) * 8 and exchange_feasible(counts, infos)
        plans = [plan_exchange(counts, r) for r in range(2)]
        bufs = []
        for r in range(2):
            _, send_slices, send_splits, _, _ = plans[r]
            buf = torch.zeros(max(sum(send_splits), 1) * W, dtype=torch.int32, device="cuda")
            ranks[r].log_export(send_slices, buf.data_ptr())
            ranks[r].stream_sync()
            bufs.append(buf)
        total = np.zeros((len(kList), 2, 65536), dtype=np.uint64)
        for r in range(2):
            owned, _, _, recv_splits, runs = plans[r]
            assert sum(recv_splits) == plans[1 - r][2][r]
            ranks[r].log_import(bufs[1 - r].data_ptr(), sum(recv_splits), runs)
            ranks[r].flush_slices(owned)
            total += ranks[r].hist_slices(owned)
            with pytest.raises(nt.NtcError):
                ranks[r].finish(counters=True)              # only the owned slices are defined now
        total[:, :, 0] = rB - total[:, :, 1:].sum(axis=2)
        assert np.array_equal(total, want_p.astype(np.uint64))
        f1 = s0.totals_nosync() + s1.totals_nosync()
        assert np.array_equal(f1, wf1)
        # a context is usable again after a reset, and a flushed log is no longer exportable
        s0.reset()
        s0.submit(words, None, n, stride)
        s0.flush()
        assert s0.log_counts()[1] is False
        t, f1b, _ = s0.finish(counters=True, hist=False)
        assert np.array_equal(t.reshape(-1), want) and np.array_equal(f1b, wf1)


@pytest.mark.parametrize("name,n,L,kList,sBits,S,mode", [
    ("config 2 (full)", 10_000_000, 150, [32], 7, 1, 0),
    ("config 3 (10M-read prefix)", 10_000_000, 150, [32, 64, 96, 128], 7, 2, 0),
    ("config 4 (10M-read prefix of GPU 0's shard)", 10_000_000, 150, [64], 11, 3, 0),
    ("config 5 (1M-read prefix, with N)", 1_000_000, 10_000, [31], 11, 4, 2),
])
def test_at_size_against_the_reference(oracle, name, n, L, kList, sBits, S, mode):
    """BASELINE configs at size (SURVEY 8d prefixes) against the UNMODIFIED reference's own ntRead (oracle/_ref, OpenMP over
    reads on the host cores; the oracle port where _ref was not built): r = 27 as in ntcard.cpp:57, the device's counter-value
    histogram must equal the bincount of the reference's uint16 sketch, table by table, and F1 must be equal."""
    import os
    from oracle.pyoracle import Reference
    rBits, threads = 27, os.cpu_count() or 1
    nK, rB = len(kList), 1 << 27
    want_p = np.zeros((nK, 2, 65536), dtype=np.int64)
    want_f1 = np.zeros(nK, dtype=np.uint64)
    ref = Reference() if Reference.available() else None
    want_sk = np.zeros(nK * 2 * rB, dtype=np.uint16)
    chunk = 2_000_000 if L <= 300 else 100_000          # reads per host chunk (ASCII in RAM: <= 1 GB)
    with nt.Sketch(kList, rBits=rBits, sBits=sBits) as sk:
        if ref is not None:
            ref.set_opts(rBits, sBits, nK)
        for c0 in range(0, n, chunk):
            nc = min(chunk, n - c0)
            reads = oracle.gen_reads(S, c0, nc, L, mode, 0)
            off = np.arange(nc + 1, dtype=np.uint64) * L
            tot = np.zeros(nK, dtype=np.uint64)
            if ref is not None:
                ref.ntread_batch(reads, off, kList, want_sk, tot, threads)
            else:
                oracle.ntread_batch(reads, off, kList, rBits, sBits, want_sk, tot, threads)
            want_f1 += tot
            if mode == 2:                                # reads with N: the documented host route (packer -> ragged batch)
                w, o = nt.pack_chars(reads, off, min_len=min(kList))
                sk.submit(w, o)
            else:                                        # uniform reads: the same reads from the device generator
                stride = nt.stride_words(L)
                sk.submit(nt.gen_packed(S, c0, nc, L, 0, 0, stride), None, nc, stride)
            del reads
        _, f1, p = sk.finish(counters=False, hist=True)
    for t in range(2 * nK):
        want_p.reshape(-1, 65536)[t] = np.bincount(want_sk[t * rB:(t + 1) * rB], minlength=65536)
    assert np.array_equal(f1, want_f1), (name, f1, want_f1)
    assert np.array_equal(p.astype(np.int64), want_p), name
    assert int(want_p[:, :, 1:].sum()) > 0


@pytest.mark.parametrize("pool", [None, "64"])
def test_fused_kernel_experiment_stays_exact(oracle, monkeypatch, pool):
    """NTC_FUSED=1: the fused scan + hash + append kernel of round 2 (fused_kernel.cuh; measured slower than scan + hit, DESIGN 5b, and
    therefore not the default) must keep producing the oracle's sketch -- uniform and mixed-length tiles, several k classes, and with a
    64-block pool (tiles deferred to the second pass, then direct increments)."""
    monkeypatch.setenv("NTC_FUSED", "1")
    if pool:
        monkeypatch.setenv("NTC_POOL_BLOCKS", pool)
    kList, rBits, sBits, L, n = [12, 25, 32, 64, 96], 20, 7, 150, 12000
    a = oracle.gen_reads(74, 0, n, L, 1, n // 6)
    reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
    reads[100:140] = [b"A" * L] * 20 + [b"ACGT" * 37 + b"AC"] * 20      # low complexity: queue overflow / re-scan rounds
    want, wf1 = oracle.sketch_reads(reads, kList, rBits, sBits, nthreads=4)
    stride = nt.stride_words(L)
    words, off = nt.pack_reads(reads, min_len=1)
    uni = np.zeros(n * stride, dtype=np.uint32)
    for i in range(n):                                                        # all reads are N-free: one record each
        uni[i * stride:i * stride + (off[i + 1] - off[i])] = words[off[i]:off[i + 1]]
    with nt.Sketch(kList, rBits=rBits, sBits=sBits) as sk:
        sk.set_kernel(nt.KERNEL_BITSLICE)
        sk.submit(uni[:5000 * stride], None, 5000, stride)
        sk.submit(uni[5000 * stride:], None, n - 5000, stride)
        t, f1, _ = sk.finish(counters=True, hist=False)
    assert np.array_equal(f1, wf1)
    assert np.array_equal(t.reshape(-1), want)


def test_headerless_uniform_batches(oracle):
    """ntc_submit_bases: reads of one length without length words (40 instead of 44 bytes per 150 bp read over PCIe).  Same sketch as
    the oracle's for lengths around the word and 16-byte-group boundaries, pinned and pageable sources, small batches (general kernel)
    and the nthll context."""
    kList, rBits, sBits = [25, 32, 64], 22, 7
    for L, n in ((150, 9000), (144, 5000), (160, 5000), (100, 300), (64, 2000)):
        a = oracle.gen_reads(73, 0, n, L, 1, max(n // 5, 1))
        reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
        want, wf1 = oracle.sketch_reads(reads, kList, rBits, sBits, nthreads=4)
        stride = nt.stride_words(L, False)
        wpr = (L + 15) // 16
        words = nt.gen_packed(73, 0, n, L, 1, max(n // 5, 1), stride)
        bases = np.ascontiguousarray(words.reshape(n, stride)[:, 1:1 + wpr]).reshape(-1)
        pin = nt.PinnedBuffer(len(bases))
        pin.array[:] = bases
        with nt.Sketch(kList, rBits=rBits, sBits=sBits) as sk:
            for src in (bases, pin.array):
                sk.reset()
                half = n // 2
                sk.submit_bases(src[:half * wpr], half, L)
                sk.submit_bases(src[half * wpr:], n - half, L)
                t, f1, _ = sk.finish(counters=True, hist=False)
                assert np.array_equal(f1, wf1), L
                assert np.array_equal(t.reshape(-1), want), L
        pin.free()
        if L == 150:
            with nt.HllSketch(32, 16) as h:
                h.submit_bases(bases, n, L)
                regs, nk = h.finish()
                assert nk == n * (L - 32 + 1)
                assert np.array_equal(regs, oracle.hll_registers(reads, 32, 16, nthreads=4))


def test_device_resident_ragged_batches(oracle):
    """Ragged batches that already live in device memory (ntc_submit_device with offsets): the library finds the longest
    record and checks the offsets ON the device, then short records are padded for the pipeline and long ones re-tiled --
    same result as the oracle either way; broken offsets are refused."""
    import torch
    kList, rBits, sBits = [31, 64], 22, 7
    for L, n in ((150, 30000), (6000, 600)):
        a = oracle.gen_reads(72, 0, n, L, 2, 0)
        reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
        want, wf1 = oracle.sketch_reads(reads, kList, rBits, sBits, nthreads=4)
        w, off = nt.pack_reads(reads, min_len=min(kList))
        dw = torch.from_numpy(w.view(np.int32)).cuda()
        do = torch.from_numpy(off.view(np.int32)).cuda()
        with nt.Sketch(kList, rBits=rBits, sBits=sBits) as sk:
            sk.submit_device(dw.data_ptr(), len(w), len(off) - 1, 0, d_off=do.data_ptr())
            t, f1, _ = sk.finish(counters=True, hist=False)
            assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want), L
            bad = off.copy()
            bad[len(bad) // 2] = bad[-1] + 7          # not monotonic / past the end
            db = torch.from_numpy(bad.view(np.int32)).cuda()
            with pytest.raises(nt.NtcError):
                sk.submit_device(dw.data_ptr(), len(w), len(off) - 1, 0, d_off=db.data_ptr())


def test_concurrent_submit_from_several_threads(oracle):
    """ntRead is called concurrently on one shared sketch (one OpenMP thread per file, ntcard.cpp:445); so may ntc_submit be:
    four host threads submit interleaved batches (ragged and uniform) to ONE context; the sketch must equal the oracle's."""
    import threading
    n, L, kList, rBits, sBits = 40000, 150, [32, 64], 22, 7
    a = oracle.gen_reads(71, 0, n, L, 2, 0)
    reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
    want, wf1 = oracle.sketch_reads(reads, kList, rBits, sBits, nthreads=4)
    errors = []
    with nt.Sketch(kList, rBits=rBits, sBits=sBits) as sk:
        def worker(t):
            try:
                for b0 in range(t * 500, n, 4 * 500):
                    sk.submit_reads(reads[b0:b0 + 500])
            except Exception as e:  # pragma: no cover
                errors.append(e)
        th = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        assert not errors, errors
        t, f1, _ = sk.finish(counters=True, hist=False)
    assert np.array_equal(f1, wf1)
    assert np.array_equal(t.reshape(-1), want)


@pytest.mark.parametrize("world", [1, 2, 3])
def test_peer_memory_reduction_between_contexts_on_one_gpu(oracle, world):
    """The multi-GPU reduction over peer memory (ntc_log_status_device / ntc_reduce_owned) exercised on ONE device: `world`
    contexts play the ranks, each sketches its share of the reads; every rank then zeroes the slices it owns (s % world ==
    rank), applies to them the hit-log entries of ALL ranks -- read straight out of the other contexts' logs -- and
    histograms them.  The summed histograms must equal the histogram of one sketch of all reads (oracle), F1 likewise."""
    import torch
    n, L, kList, rBits, sBits = 30720, 150, [32, 64], 25, 7
    stride = nt.stride_words(L)
    words = nt.gen_packed(62, 0, n, L, 1, n // 4, stride)
    a = oracle.gen_reads(62, 0, n, L, 1, n // 4)
    reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
    want, wf1 = oracle.sketch_reads(reads, kList, rBits, sBits, nthreads=4)
    rB = 1 << rBits
    want_p = np.stack([np.bincount(want[t * rB:(t + 1) * rB], minlength=65536) for t in range(2 * len(kList))]).reshape(len(kList), 2, 65536)
    nK = len(kList)
    ranks = [nt.Sketch(kList, rBits=rBits, sBits=sBits) for _ in range(world)]
    try:
        for r, sk in enumerate(ranks):
            sk.peer_attach_contexts(r, ranks)
        for rep in range(2):                                   # twice: a context is usable again after reset
            status = [torch.zeros(nK + 1, dtype=torch.int64, device="cuda") for _ in ranks]
            hist = [torch.zeros(nK * 2 * 65536, dtype=torch.int32, device="cuda") for _ in ranks]
            per = (n + world - 1) // world
            for r, sk in enumerate(ranks):
                sk.reset()
                lo, hi = r * per, min(n, (r + 1) * per)
                sk.submit(words[lo * stride:hi * stride], None, hi - lo, stride)
                sk.log_status_device(status[r].data_ptr())
            for sk in ranks:
                sk.stream_sync()                               # stands in for the status all-reduce (sum + barrier)
            tot = torch.stack(status).sum(dim=0)
            assert int(tot[nK]) == 0
            assert np.array_equal(tot[:nK].cpu().numpy().astype(np.uint64), wf1)
            for r, sk in enumerate(ranks):
                sk.reduce_owned(tot.data_ptr(), hist[r].data_ptr())
            for sk in ranks:
                sk.stream_sync()                               # stands in for the histogram all-reduce
            p = torch.stack(hist).cpu().numpy().view(np.uint32).astype(np.uint64).sum(axis=0).reshape(nK, 2, 65536)
            p[:, :, 0] = rB - p[:, :, 1:].sum(axis=2)
            assert np.array_equal(p, want_p.astype(np.uint64))
            with pytest.raises(nt.NtcError):
                ranks[0].finish(counters=True)                 # only the owned slices are defined now
        # an incomplete log (flushed) raises the flag and the kernels leave the histogram untouched
        ranks[0].reset()
        ranks[0].submit(words, None, n, stride)
        ranks[0].flush()
        st = torch.zeros(nK + 1, dtype=torch.int64, device="cuda")
        ranks[0].log_status_device(st.data_ptr())
        ranks[0].stream_sync()
        assert int(st[nK]) == 1
    finally:
        for sk in ranks:
            sk.close()


def test_tightly_packed_uniform_batches_take_the_pipeline(oracle):
    """Uniform-stride batches whose stride is not a multiple of 4 words (44 bytes per 150 bp read) are padded on the
    device and hashed by the pipeline; forced NTC_KERNEL_BITSLICE must accept them and the result stays exact."""
    for L, kList in ((150, [32, 64]), (120, [31]), (171, [12, 32])):
        n = 5000
        stride = nt.stride_words(L, False)
        assert stride % 4 != 0 or L == 171
        words, _, want, wf1 = _uniform_case(oracle, 83, n, L, kList, 19, 7)       # aligned copy, for the oracle result
        tight = nt.gen_packed(83, 0, n, L, 1, max(1, n // 4), stride)
        with nt.Sketch(kList, rBits=19, sBits=7) as sk:
            sk.set_kernel(nt.KERNEL_BITSLICE)
            sk.submit(tight[:2000 * stride], None, 2000, stride)
            sk.submit(tight[2000 * stride:], None, n - 2000, stride)
            t, f1, _ = sk.finish(counters=True, hist=False)
        assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want), L


def test_long_uniform_stride_records_are_retiled(oracle):
    """Long reads at a uniform stride (device generators, fixed-length long-read batches) get explicit offsets and go
    through the same re-tiling as ragged long records."""
    n, L, kList = 1500, 3000, [32, 64]
    words, stride, want, wf1 = _uniform_case(oracle, 91, n, L, kList, 20, 7, mode=0)
    for kernel in (nt.KERNEL_AUTO, nt.KERNEL_ROLL64):
        with nt.Sketch(kList, rBits=20, sBits=7) as sk:
            sk.set_kernel(kernel)
            sk.submit(words, None, n, stride)
            t, f1, _ = sk.finish(counters=True, hist=False)
        assert np.array_equal(f1, wf1) and np.array_equal(t.reshape(-1), want), kernel
