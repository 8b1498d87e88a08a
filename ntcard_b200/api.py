"""ctypes binding of libntcard_b200.so (include/ntcard_b200.h) -- the host-side mirror of the
reference's ntRead / ntComp / compEst boundary (ntcard.cpp:132-158, 237-275).

There is no Python or CPU implementation of the hot path in this package: if the CUDA library is
missing, importing this module raises; if no B200 is visible, `Sketch(...)` raises NtcError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libntcard_b200.so")

NTC_OK, NTC_EINVAL, NTC_ENODEVICE, NTC_ECUDA, NTC_ENOMEM, NTC_ESTATE = 0, -1, -2, -3, -4, -5
NTC_MAX_K = 16
KERNEL_AUTO, KERNEL_ROLL64, KERNEL_BITSLICE = 0, 1, 2

# every symbol include/ntcard_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "ntc_create", "ntc_destroy", "ntc_reset", "ntc_set_kernel", "ntc_set_gap", "ntc_submit", "ntc_submit_bases", "ntc_submit_device", "ntc_wait",
    "ntc_sync", "ntc_flush", "ntc_log_info", "ntc_log_counts", "ntc_log_export", "ntc_log_import", "ntc_flush_slices",
    "ntc_hist_slices", "ntc_stream_sync", "ntc_counters_device", "ntc_hist_range", "ntc_totals", "ntc_totals_nosync", "ntc_set_totals", "ntc_finish", "ntc_estimate",
    "ntc_host_alloc", "ntc_host_free", "ntc_pack_bound", "ntc_pack_bound_k", "ntc_pack_seqs", "ntc_gen_ascii", "ntc_gen_packed",
    "ntc_gen_packed_device", "ntc_stride_words", "ntc_stats", "ntc_kernel_time", "ntc_stage_times", "ntc_device_count",
    "ntc_last_error", "ntc_version",
    "ntc_hll_create", "ntc_hll_registers_device", "ntc_hll_finish", "ntc_hll_estimate", "ntc_check_offsets",
    "ntc_peer_export", "ntc_peer_attach", "ntc_peer_attach_contexts", "ntc_log_status_device", "ntc_reduce_owned",
]
PEER_HANDLE_BYTES = 128


class NtcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ntcard_b200 error {code}: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` or `make` "
            "at the repository root. ntcard_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u32p, u64p, u16p, dblp = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint16), C.POINTER(C.c_double)
    sig = {
        "ntc_create": (C.c_int, [C.POINTER(vp), u32p, C.c_uint, C.c_uint, C.c_uint, C.c_int, vp, vp]),
        "ntc_destroy": (None, [vp]),
        "ntc_reset": (C.c_int, [vp]),
        "ntc_set_kernel": (C.c_int, [vp, C.c_int]),
        "ntc_set_gap": (C.c_int, [vp, C.c_uint]),
        "ntc_submit": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_uint32, u64p]),
        "ntc_submit_bases": (C.c_int, [vp, vp, C.c_size_t, C.c_uint32, u64p]),
        "ntc_submit_device": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_size_t, C.c_uint32]),
        "ntc_wait": (C.c_int, [vp, C.c_uint64]),
        "ntc_sync": (C.c_int, [vp]),
        "ntc_flush": (C.c_int, [vp]),
        "ntc_log_info": (C.c_int, [vp, u32p, u64p, u32p]),
        "ntc_log_counts": (C.c_int, [vp, vp, C.POINTER(C.c_int), vp]),
        "ntc_stream_sync": (C.c_int, [vp]),
        "ntc_log_export": (C.c_int, [vp, vp, C.c_uint32, vp]),
        "ntc_log_import": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32]),
        "ntc_flush_slices": (C.c_int, [vp, vp]),
        "ntc_hist_slices": (C.c_int, [vp, vp, vp, vp]),
        "ntc_counters_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "ntc_totals": (C.c_int, [vp, u64p]),
        "ntc_totals_nosync": (C.c_int, [vp, u64p]),
        "ntc_set_totals": (C.c_int, [vp, u64p]),
        "ntc_finish": (C.c_int, [vp, vp, u64p, vp]),
        "ntc_hist_range": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, vp]),
        "ntc_estimate": (C.c_int, [vp, vp, C.c_uint, C.c_uint, C.c_uint, dblp, dblp]),
        "ntc_host_alloc": (vp, [C.c_size_t]),
        "ntc_host_free": (None, [vp]),
        "ntc_pack_bound": (C.c_size_t, [C.c_size_t, C.c_size_t]),
        "ntc_pack_bound_k": (C.c_size_t, [C.c_size_t, C.c_size_t, C.c_uint32]),
        "ntc_pack_seqs": (C.c_int, [vp, u64p, C.c_size_t, C.c_uint32, vp, C.c_size_t, C.POINTER(C.c_size_t), vp,
                                    C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "ntc_gen_ascii": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint64, vp]),
        "ntc_gen_packed": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint64, C.c_uint32, vp]),
        "ntc_gen_packed_device": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_uint64,
                                            C.c_uint32, vp]),
        "ntc_stride_words": (C.c_uint32, [C.c_uint32, C.c_int]),
        "ntc_stats": (C.c_int, [vp, u64p, u64p]),
        "ntc_kernel_time": (C.c_int, [vp, dblp, u64p]),
        "ntc_stage_times": (C.c_int, [vp, dblp]),
        "ntc_device_count": (C.c_int, []),
        "ntc_last_error": (C.c_char_p, []),
        "ntc_version": (C.c_char_p, []),
        "ntc_hll_create": (C.c_int, [C.POINTER(vp), C.c_uint, C.c_uint, C.c_int, vp, vp]),
        "ntc_hll_registers_device": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "ntc_hll_finish": (C.c_int, [vp, vp, u64p]),
        "ntc_hll_estimate": (C.c_int, [vp, C.c_uint, C.c_int, dblp]),
        "ntc_check_offsets": (C.c_int, [vp, C.c_size_t, C.c_size_t, u32p]),
        "ntc_peer_export": (C.c_int, [vp, vp]),
        "ntc_peer_attach": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "ntc_peer_attach_contexts": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "ntc_log_status_device": (C.c_int, [vp, vp]),
        "ntc_reduce_owned": (C.c_int, [vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def _check(rc):
    if rc != NTC_OK:
        raise NtcError(rc, lib.ntc_last_error().decode("utf-8", "replace"))


def device_count():
    return lib.ntc_device_count()


def stride_words(L, align4=True):
    return lib.ntc_stride_words(L, 1 if align4 else 0)


def apply_sbits_rule(total_input_bytes, sBits=11):
    """ntcard.cpp:427-431: sBits is forced to 7 when the input files total < 50 GB."""
    return 7 if total_input_bytes < 50000000000 else sBits


# ---- host helpers -------------------------------------------------------------------------------
def pack_reads(reads, min_len=1):
    """Split each read (bytes) at characters outside ACGTUacgtu and 2-bit pack the segments.
    Returns (words uint32[n_words], off uint32[n_rec+1])."""
    lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
    soff = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum(lens, out=soff[1:])
    chars = np.frombuffer(b"".join(reads) + b"\0", dtype=np.uint8)
    return pack_chars(chars, soff, min_len)


def pack_chars(chars, seq_off, min_len=1):
    """chars: uint8 array of concatenated sequences; seq_off: uint64[n+1]."""
    chars = np.ascontiguousarray(chars, dtype=np.uint8)
    seq_off = np.ascontiguousarray(seq_off, dtype=np.uint64)
    n = len(seq_off) - 1
    total = int(seq_off[-1])
    cap_w = lib.ntc_pack_bound_k(n, total, max(1, min_len))
    cap_r = n + total // (max(1, min_len) + 1) + 2
    words = np.empty(cap_w, dtype=np.uint32)
    off = np.empty(cap_r + 1, dtype=np.uint32)
    nw, nr, cons = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    _check(lib.ntc_pack_seqs(chars.ctypes.data, seq_off.ctypes.data_as(C.POINTER(C.c_uint64)), n, min_len,
                             words.ctypes.data, cap_w, C.byref(nw), off.ctypes.data, cap_r, C.byref(nr), C.byref(cons)))
    return words[:nw.value].copy(), off[:nr.value + 1].copy()


def check_offsets(off, n_words):
    """ntc_submit's checks of a ragged batch's offsets (uint32[n_rec+1]); returns the word count of the longest record."""
    off = np.ascontiguousarray(off, dtype=np.uint32)
    mx = C.c_uint32()
    _check(lib.ntc_check_offsets(off.ctypes.data, len(off) - 1, n_words, C.byref(mx)))
    return mx.value


def gen_ascii(seed, first, n, L, mode=0, U=0):
    out = np.empty(n * L, dtype=np.uint8)
    _check(lib.ntc_gen_ascii(seed, first, n, L, mode, U, out.ctypes.data))
    return out


def gen_packed(seed, first, n, L, mode=0, U=0, stride=None, out=None):
    stride = stride or stride_words(L)
    if out is None:
        out = np.empty(n * stride, dtype=np.uint32)
    _check(lib.ntc_gen_packed(seed, first, n, L, mode, U, stride, out.ctypes.data))
    return out


def estimate(p_hist=None, t_counter=None, rBits=27, sBits=7, covMax=1000):
    """compEst (ntcard.cpp:237-275).  Returns (F0, f[0..covMax]) with f[0] = 0."""
    F0 = C.c_double()
    covMax = min(covMax, 65535)
    f = np.zeros(covMax + 1, dtype=np.float64)
    p = np.ascontiguousarray(p_hist, dtype=np.uint32).ctypes.data if p_hist is not None else None
    t = np.ascontiguousarray(t_counter, dtype=np.uint16).ctypes.data if t_counter is not None else None
    _check(lib.ntc_estimate(p, t, rBits, sBits, covMax, C.byref(F0), f.ctypes.data_as(C.POINTER(C.c_double))))
    return F0.value, f


class PinnedBuffer:
    """Pinned host memory from ntc_host_alloc, viewed as a numpy uint32 array."""

    def __init__(self, n_words):
        self.ptr = lib.ntc_host_alloc(n_words * 4)
        if not self.ptr:
            raise NtcError(NTC_ENOMEM, lib.ntc_last_error().decode())
        self.n_words = n_words
        self.array = np.ctypeslib.as_array((C.c_uint32 * n_words).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            lib.ntc_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Sketch:
    """One ntCard sketch on one GPU: the reference's t_Counter + totalKmers for a k list.

    sketch = Sketch([32], rBits=27, sBits=7)       # ntcard.cpp:437-439
    sketch.submit(words, off)                      # == ntRead over every record (ntcard.cpp:147-158)
    t, F1, p = sketch.finish(counters=True)        # uint16 sketch, totKmer, counter-value histogram
    F0, f = estimate(p_hist=p[0], rBits=27, sBits=7, covMax=1000)
    """

    def __init__(self, kList, rBits=27, sBits=7, device=0, d_counters=None, stream=None):
        self.kList = [int(k) for k in kList]
        self.nK, self.rBits, self.sBits, self.device = len(self.kList), rBits, sBits, device
        kl = (C.c_uint * self.nK)(*self.kList)
        h = C.c_void_p()
        _check(lib.ntc_create(C.byref(h), C.cast(kl, C.POINTER(C.c_uint32)), self.nK, rBits, sBits, device,
                              C.c_void_p(d_counters) if d_counters else None,
                              C.c_void_p(stream) if stream else None))
        self.h = h
        self.stream_handle = int(stream) if stream else None   # the CUDA stream all device work is ordered on (None: own)

    def close(self):
        if getattr(self, "h", None):
            lib.ntc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def reset(self):
        _check(lib.ntc_reset(self.h))

    def set_kernel(self, kernel):
        _check(lib.ntc_set_kernel(self.h, kernel))

    def set_gap(self, gap):
        """The reference's -g: spaced seed with `gap` don't-care bases in the middle (one k)."""
        _check(lib.ntc_set_gap(self.h, gap))

    def submit(self, words, off=None, n_rec=None, stride=0):
        """Host batch.  words: uint32 array (numpy or PinnedBuffer.array slice); off: uint32[n_rec+1] or
        None for a uniform batch of n_rec records of `stride` words.  Returns the ticket."""
        words = np.ascontiguousarray(words, dtype=np.uint32)
        t = C.c_uint64()
        if off is not None:
            off = np.ascontiguousarray(off, dtype=np.uint32)
            n_rec = len(off) - 1
            _check(lib.ntc_submit(self.h, words.ctypes.data, len(words), off.ctypes.data, n_rec, 0, C.byref(t)))
        else:
            _check(lib.ntc_submit(self.h, words.ctypes.data, len(words), None, n_rec, stride, C.byref(t)))
        return t.value

    def submit_bases(self, bases, n_rec, L):
        """Host batch of n_rec reads of ONE length L without length words: ceil(L/16) uint32 words of bases per read."""
        bases = np.ascontiguousarray(bases, dtype=np.uint32)
        assert len(bases) == n_rec * ((L + 15) // 16)
        t = C.c_uint64()
        _check(lib.ntc_submit_bases(self.h, bases.ctypes.data, n_rec, L, C.byref(t)))
        return t.value

    def submit_device(self, d_words, n_words, n_rec, stride=0, d_off=None):
        _check(lib.ntc_submit_device(self.h, d_words, n_words, d_off, n_rec, stride))

    def submit_reads(self, reads):
        """Convenience: ASCII reads (list of bytes) -> N-split -> pack -> submit."""
        words, off = pack_reads(reads, min_len=min(self.kList))
        if len(off) > 1:
            self.submit(words, off)

    def wait(self, ticket):
        _check(lib.ntc_wait(self.h, ticket))

    def sync(self):
        _check(lib.ntc_sync(self.h))

    def flush(self):
        """Apply the pending sketch increments to the counters in HBM (asynchronous, stream ordered)."""
        _check(lib.ntc_flush(self.h))

    # ---- hit-log exchange (multi-GPU sparse reduction; include/ntcard_b200.h) ----
    def log_info(self):
        n, cps, epb = C.c_uint32(), C.c_uint64(), C.c_uint32()
        _check(lib.ntc_log_info(self.h, C.byref(n), C.byref(cps), C.byref(epb)))
        return n.value, cps.value, epb.value

    def log_counts(self):
        """(nblk uint32[n_slices], exportable bool, (blocks used, blocks in pool, list capacity per slice))"""
        n = self.log_info()[0]
        nblk = np.zeros(n, dtype=np.uint32)
        info = np.zeros(3, dtype=np.uint32)
        ok = C.c_int()
        _check(lib.ntc_log_counts(self.h, nblk.ctypes.data, C.byref(ok), info.ctypes.data))
        return nblk, bool(ok.value), tuple(int(x) for x in info)

    def stream_sync(self):
        _check(lib.ntc_stream_sync(self.h))

    WIRE_BLOCK_WORDS = 260  # a log block on the wire: 256 entries, fill, padding

    def log_export(self, slices, d_blocks):
        sl = np.ascontiguousarray(slices, dtype=np.uint32)
        _check(lib.ntc_log_export(self.h, sl.ctypes.data, len(sl), d_blocks))

    def log_import(self, d_blocks, n_blocks, runs):
        r = np.ascontiguousarray(runs, dtype=np.uint32).reshape(-1, 2)
        _check(lib.ntc_log_import(self.h, d_blocks, n_blocks, r.ctypes.data, len(r)))

    def flush_slices(self, owned):
        o = np.ascontiguousarray(owned, dtype=np.uint8)
        _check(lib.ntc_flush_slices(self.h, o.ctypes.data))

    def hist_slices(self, owned, d_out=None):
        """Histogram (v >= 1) of the owned slices: to a device buffer (d_out: pointer, asynchronous) or as numpy."""
        o = np.ascontiguousarray(owned, dtype=np.uint8)
        if d_out is not None:
            _check(lib.ntc_hist_slices(self.h, o.ctypes.data, None, d_out))
            return None
        p = np.zeros((self.nK, 2, 65536), dtype=np.uint32)
        _check(lib.ntc_hist_slices(self.h, o.ctypes.data, p.ctypes.data, None))
        return p

    # ---- multi-GPU reduction over peer memory (include/ntcard_b200.h) ----
    def peer_export(self):
        """IPC handles of this context's hit log (bytes, PEER_HANDLE_BYTES long) for the other ranks of the node."""
        buf = np.zeros(PEER_HANDLE_BYTES, dtype=np.uint8)
        _check(lib.ntc_peer_export(self.h, buf.ctypes.data))
        return buf

    def peer_attach(self, world, rank, all_handles):
        """Map the hit logs of all ranks (all_handles: uint8 [world, PEER_HANDLE_BYTES], row r = rank r's peer_export())."""
        hs = np.ascontiguousarray(all_handles, dtype=np.uint8).reshape(world, PEER_HANDLE_BYTES)
        _check(lib.ntc_peer_attach(self.h, world, rank, hs.ctypes.data))
        self.peer_world, self.peer_rank = world, rank

    def peer_attach_contexts(self, rank, sketches):
        """The same for Sketch objects of this process (sketches[rank] is self)."""
        arr = (C.c_void_p * len(sketches))(*[x.h for x in sketches])
        _check(lib.ntc_peer_attach_contexts(self.h, len(sketches), rank, arr))
        self.peer_world, self.peer_rank = len(sketches), rank

    def log_status_device(self, d_status):
        """d_status (device pointer, int64 [nK + 1]) := F1 per k, 1 if this rank's log is incomplete (stream ordered)."""
        _check(lib.ntc_log_status_device(self.h, d_status))

    def reduce_owned(self, d_status, d_p_hist):
        """Apply every rank's log entries of the slices this rank owns, histogram them into d_p_hist (device, uint32
        [nK, 2, 65536], values >= 1); stream ordered, no host synchronisation.  d_status: the all-reduced status."""
        _check(lib.ntc_reduce_owned(self.h, d_status, d_p_hist))

    def counters_device(self):
        p, n = C.c_void_p(), C.c_size_t()
        _check(lib.ntc_counters_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def totals(self):
        t = np.zeros(self.nK, dtype=np.uint64)
        _check(lib.ntc_totals(self.h, t.ctypes.data_as(C.POINTER(C.c_uint64))))
        return t

    def totals_nosync(self):
        """F1 as counted so far on the device, without flushing (valid after a stream_sync / log_counts)."""
        t = np.zeros(self.nK, dtype=np.uint64)
        _check(lib.ntc_totals_nosync(self.h, t.ctypes.data_as(C.POINTER(C.c_uint64))))
        return t

    def set_totals(self, tot):
        t = np.ascontiguousarray(tot, dtype=np.uint64)
        _check(lib.ntc_set_totals(self.h, t.ctypes.data_as(C.POINTER(C.c_uint64))))

    def finish(self, counters=False, hist=True):
        """Returns (t_Counter uint16 [nK,2,2^r] or None, totKmer uint64 [nK], p_hist uint32 [nK,2,65536] or None)."""
        t = np.empty((self.nK, 2, 1 << self.rBits), dtype=np.uint16) if counters else None
        p = np.empty((self.nK, 2, 65536), dtype=np.uint32) if hist else None
        tot = np.zeros(self.nK, dtype=np.uint64)
        _check(lib.ntc_finish(self.h, t.ctypes.data if counters else None, tot.ctypes.data_as(C.POINTER(C.c_uint64)),
                              p.ctypes.data if hist else None))
        return t, tot, p

    def hist_range(self, d_counters, first, n):
        """Counter-value histogram (v >= 1) of the slice [first, first+n) of the flat counters held at the
        device pointer d_counters (e.g. this rank's reduce-scatter output).  Returns uint32 [nK,2,65536]."""
        p = np.zeros((self.nK, 2, 65536), dtype=np.uint32)
        _check(lib.ntc_hist_range(self.h, d_counters, first, n, p.ctypes.data))
        return p

    def gen_packed_device(self, seed, first, n, L, mode, U, stride, d_words):
        _check(lib.ntc_gen_packed_device(self.h, seed, first, n, L, mode, U, stride, d_words))

    def stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(lib.ntc_stats(self.h, C.byref(a), C.byref(b)))
        return {"launches": a.value, "batches": b.value}

    def stage_times(self):
        """(scan_ms, hit_ms, apply_ms) accumulated since the last call; call after sync()."""
        ms = (C.c_double * 3)()
        _check(lib.ntc_stage_times(self.h, C.cast(ms, C.POINTER(C.c_double))))
        return tuple(ms)

    def kernel_time(self):
        ms, n = C.c_double(), C.c_uint64()
        _check(lib.ntc_kernel_time(self.h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def hll_estimate(regs, nBits=16, canon=True):
    """The estimate nthll prints (nthll.cpp:243-254) from uint8 registers; host arithmetic, no device."""
    regs = np.ascontiguousarray(regs, dtype=np.uint8)
    if len(regs) != 1 << nBits:
        raise ValueError("hll_estimate: need 2^nBits registers")
    e = C.c_double()
    _check(lib.ntc_hll_estimate(regs.ctypes.data, nBits, 1 if canon else 0, C.byref(e)))
    return e.value


class HllSketch(Sketch):
    """nthll's HyperLogLog registers on one GPU (nthll.cpp:92-104): a context in nthll mode.  Batches go in through the
    same submit / submit_device / submit_reads / wait / sync / reset as Sketch; the ntCard-sketch calls raise NtcError.

    h = HllSketch(k=64, nBits=16)
    h.submit_reads(reads)
    regs, n_kmers = h.finish()              # uint8 [2^nBits], k-mers hashed
    F0 = hll_estimate(regs, 16)             # nthll prints int(F0)
    """

    def __init__(self, k=64, nBits=16, device=0, d_regs=None, stream=None):
        self.kList, self.nK, self.rBits, self.sBits, self.device, self.nBits = [int(k)], 1, 0, 0, device, nBits
        h = C.c_void_p()
        _check(lib.ntc_hll_create(C.byref(h), int(k), nBits, device, C.c_void_p(d_regs) if d_regs else None,
                                  C.c_void_p(stream) if stream else None))
        self.h = h
        self.stream_handle = int(stream) if stream else None

    def registers_device(self):
        p, n = C.c_void_p(), C.c_size_t()
        _check(lib.ntc_hll_registers_device(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def finish(self, registers=True):
        """Returns (registers uint8 [2^nBits] or None, number of k-mers hashed)."""
        regs = np.empty(1 << self.nBits, dtype=np.uint8) if registers else None
        tot = np.zeros(1, dtype=np.uint64)
        _check(lib.ntc_hll_finish(self.h, regs.ctypes.data if registers else None, tot.ctypes.data_as(C.POINTER(C.c_uint64))))
        return regs, int(tot[0])


def write_hist(path, F1, F0, f, covMax):
    """outDefault's per-k file (ntcard.cpp:291-294): F1, F0, then rows 1..covMax, tab separated."""
    with open(path, "w") as fp:
        fp.write(f"F1\t{int(F1)}\nF0\t{int(np.uint64(F0))}\n")
        for i in range(1, covMax + 1):
            fp.write(f"{i}\t{int(np.uint64(f[i]))}\n")
