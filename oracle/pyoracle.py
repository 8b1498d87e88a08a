"""ctypes access to the CHECKERS: oracle/libntcard_oracle.so (our C restatement)
and, when built, oracle/_ref/libntcard_ref.so (the unmodified reference behind
extern "C" shims).  TEST INFRASTRUCTURE ONLY -- imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs,
never by the ntcard_b200 package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libntcard_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libntcard_ref.so")
REF_CLI = os.path.join(HERE, "_ref", "ntcard_ref")
HLL_REF_SO = os.path.join(HERE, "_ref", "libnthll_ref.so")
HLL_REF_CLI = os.path.join(HERE, "_ref", "nthll_ref")

_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_u16p = C.POINTER(C.c_uint16)
_dblp = C.POINTER(C.c_double)


def build(ref=True):
    """Compile the oracle (and the reference shims when /root/reference exists)."""
    subprocess.run(["make", "-C", HERE, "all"] + (["ref"] if ref else []), check=True,
                   stdout=subprocess.DEVNULL)


def _ptr(a, t):
    return a.ctypes.data_as(t)


class _Lib:
    """Common surface of the oracle ('orc_') and the reference shims ('ref_')."""

    def __init__(self, path, prefix):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L = self.lib
        g = lambda n: getattr(L, prefix + n)
        g("srol").restype = C.c_uint64
        g("srol").argtypes = [C.c_uint64]
        g("sror").restype = C.c_uint64
        g("sror").argtypes = [C.c_uint64]
        g("seed").restype = C.c_uint64
        g("seed").argtypes = [C.c_ubyte]
        g("kmer_hashes").argtypes = [C.c_char_p, C.c_uint, _u64p, _u64p]
        g("hash_seq").restype = C.c_size_t
        g("hash_seq").argtypes = [C.c_char_p, C.c_size_t, C.c_uint, _u64p, _u32p, C.c_size_t]

    def srol(self, v):
        return getattr(self.lib, self.prefix + "srol")(v)

    def sror(self, v):
        return getattr(self.lib, self.prefix + "sror")(v)

    def seed(self, c):
        return getattr(self.lib, self.prefix + "seed")(c)

    def kmer_hashes(self, kmer: bytes, k=None):
        k = len(kmer) if k is None else k
        fh, rh = C.c_uint64(), C.c_uint64()
        getattr(self.lib, self.prefix + "kmer_hashes")(kmer, k, C.byref(fh), C.byref(rh))
        return fh.value, rh.value

    def hash_seq(self, seq: bytes, k):
        cap = max(len(seq), 1)
        h = np.zeros(cap, dtype=np.uint64)
        p = np.zeros(cap, dtype=np.uint32)
        n = getattr(self.lib, self.prefix + "hash_seq")(seq, len(seq), k, _ptr(h, _u64p), _ptr(p, _u32p), cap)
        return h[:n].copy(), p[:n].copy()


class Oracle(_Lib):
    def __init__(self, path=ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        super().__init__(path, "orc_")
        L = self.lib
        L.orc_srol_n.restype = C.c_uint64
        L.orc_srol_n.argtypes = [C.c_uint64, C.c_uint]
        L.orc_sample_table.restype = C.c_uint
        L.orc_sample_table.argtypes = [C.c_uint64, C.c_uint]
        L.orc_ntread.argtypes = [C.c_char_p, C.c_size_t, _u32p, C.c_uint, C.c_uint, C.c_uint, _u16p, _u64p]
        L.orc_ntread_batch.argtypes = [C.c_void_p, _u64p, C.c_size_t, _u32p, C.c_uint, C.c_uint, C.c_uint,
                                       _u16p, _u64p, C.c_int]
        L.orc_compest.argtypes = [_u16p, _u32p, C.c_uint, C.c_uint, C.c_uint, _dblp, _dblp]
        L.orc_write_hist.argtypes = [C.c_char_p, C.c_uint64, C.c_double, _dblp, C.c_uint]
        L.orc_mix64.restype = C.c_uint64
        L.orc_mix64.argtypes = [C.c_uint64]
        L.orc_gen_read.argtypes = [C.c_uint64, C.c_uint64, C.c_uint, C.c_int, C.c_uint64, C.c_char_p]
        L.orc_gen_reads.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint, C.c_int, C.c_uint64, C.c_void_p]
        L.orc_table_digest.restype = C.c_uint64
        L.orc_table_digest.argtypes = [_u16p, C.c_uint64, _u64p, _u64p, _u32p, _u64p]
        L.orc_max_threads.restype = C.c_int
        L.orc_hll_batch.argtypes = [C.c_void_p, _u64p, C.c_size_t, C.c_uint, C.c_uint, C.c_void_p, C.c_int]
        L.orc_hll_estimate.restype = C.c_double
        L.orc_hll_estimate.argtypes = [C.c_void_p, C.c_uint]
        L.orc_st_hash_seq.restype = C.c_size_t
        L.orc_st_hash_seq.argtypes = [C.c_char_p, C.c_size_t, C.c_uint, C.c_uint, _u64p, C.c_size_t]
        L.orc_stread_batch.argtypes = [C.c_void_p, _u64p, C.c_size_t, C.c_uint, C.c_uint, C.c_uint, C.c_uint, _u16p, _u64p, C.c_int]

    def srol_n(self, v, n):
        return self.lib.orc_srol_n(v, n)

    def sample_table(self, h, sBits):
        return self.lib.orc_sample_table(int(h), sBits)

    def new_sketch(self, nK, rBits):
        return np.zeros(nK * 2 * (1 << rBits), dtype=np.uint16)

    def ntread(self, seq: bytes, kList, rBits, sBits, sketch, tot):
        kl = np.asarray(kList, dtype=np.uint32)
        self.lib.orc_ntread(seq, len(seq), _ptr(kl, _u32p), len(kl), rBits, sBits, _ptr(sketch, _u16p), _ptr(tot, _u64p))

    def ntread_batch(self, seqs: np.ndarray, off: np.ndarray, kList, rBits, sBits, sketch, tot, nthreads=1):
        """seqs: uint8 array of concatenated sequences; off: uint64 [n+1]."""
        kl = np.asarray(kList, dtype=np.uint32)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        self.lib.orc_ntread_batch(seqs.ctypes.data, _ptr(off, _u64p), len(off) - 1, _ptr(kl, _u32p), len(kl),
                                  rBits, sBits, _ptr(sketch, _u16p), _ptr(tot, _u64p), nthreads)

    def sketch_reads(self, reads, kList, rBits, sBits, nthreads=1):
        """reads: list of bytes.  Returns (sketch uint16 [nK*2*2^r], totKmer uint64 [nK])."""
        sk = self.new_sketch(len(kList), rBits)
        tot = np.zeros(len(kList), dtype=np.uint64)
        lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        buf = np.frombuffer(b"".join(reads) + b"\0", dtype=np.uint8)
        self.ntread_batch(buf, off, kList, rBits, sBits, sk, tot, nthreads)
        return sk, tot

    def _flat(self, reads):
        lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        return np.frombuffer(b"".join(reads) + b"\0", dtype=np.uint8), off

    def hll_registers(self, reads, k, nBits=16, nthreads=1):
        """nthll's HyperLogLog registers (uint8 [2^nBits]) over reads (nthll.cpp:92-104, 213-241)."""
        regs = np.zeros(1 << nBits, dtype=np.uint8)
        buf, off = self._flat(reads)
        self.lib.orc_hll_batch(buf.ctypes.data, _ptr(off, _u64p), len(off) - 1, k, nBits, regs.ctypes.data, nthreads)
        return regs

    def hll_estimate(self, regs, nBits=16):
        """The estimate nthll prints (nthll.cpp:243-254)."""
        regs = np.ascontiguousarray(regs, dtype=np.uint8)
        return self.lib.orc_hll_estimate(regs.ctypes.data, nBits)

    def st_hash_seq(self, seq: bytes, k, gap):
        """Gap-seed hashes (stHashIterator + NTMSM64) of a sequence, iterator order."""
        cap = max(len(seq), 1)
        h = np.zeros(cap, dtype=np.uint64)
        self.lib.orc_st_hash_seq.restype = C.c_size_t
        n = self.lib.orc_st_hash_seq(seq, C.c_size_t(len(seq)), C.c_uint(k), C.c_uint(gap), _ptr(h, _u64p), C.c_size_t(cap))
        return h[:n].copy()

    def sketch_reads_gap(self, reads, k, gap, rBits, sBits, nthreads=1):
        """stRead (ntcard.cpp:160-171) over reads with the -g seed.  Returns (sketch uint16 [2*2^r], totKmer uint64 [1])."""
        sk = self.new_sketch(1, rBits)
        tot = np.zeros(1, dtype=np.uint64)
        buf, off = self._flat(reads)
        self.lib.orc_stread_batch(buf.ctypes.data, _ptr(off, _u64p), C.c_size_t(len(off) - 1), C.c_uint(k), C.c_uint(gap),
                                  C.c_uint(rBits), C.c_uint(sBits), _ptr(sk, _u16p), _ptr(tot, _u64p), C.c_int(nthreads))
        return sk, tot

    def compest(self, sketch=None, p_hist=None, rBits=27, sBits=7, imax=65535):
        F0 = C.c_double()
        f = np.zeros(65536, dtype=np.float64)
        t = _ptr(sketch, _u16p) if sketch is not None else None
        p = _ptr(np.ascontiguousarray(p_hist, dtype=np.uint32), _u32p) if p_hist is not None else None
        self.lib.orc_compest(t, p, rBits, sBits, imax, C.byref(F0), _ptr(f, _dblp))
        return F0.value, f

    def write_hist(self, path, F1, F0, f, covMax):
        return self.lib.orc_write_hist(path.encode(), int(F1), float(F0), _ptr(f, _dblp), covMax)

    def gen_read(self, S, i, L, mode=0, U=0):
        buf = C.create_string_buffer(L)
        self.lib.orc_gen_read(S, i, L, mode, U, buf)
        return buf.raw[:L]

    def gen_reads(self, S, first, n, L, mode=0, U=0):
        """n reads of length L as one uint8 array [n*L] (ASCII)."""
        out = np.empty(n * L, dtype=np.uint8)
        self.lib.orc_gen_reads(S, first, n, L, mode, U, out.ctypes.data)
        return out

    def table_digest(self, table: np.ndarray):
        nnz, s, first = C.c_uint64(), C.c_uint64(), C.c_uint64()
        mx = C.c_uint32()
        d = self.lib.orc_table_digest(_ptr(table, _u16p), len(table), C.byref(nnz), C.byref(s), C.byref(mx), C.byref(first))
        return {"digest": d, "nnz": nnz.value, "sum": s.value, "max": mx.value, "first": first.value}

    def max_threads(self):
        return self.lib.orc_max_threads()


class Reference(_Lib):
    """The unmodified reference (oracle/_ref/libntcard_ref.so).  NOTE: the
    reference keeps its parameters in process-wide globals (ntcard.cpp:52-67);
    set_opts() must precede ntread/compest."""

    def __init__(self, path=REF_SO):
        super().__init__(path, "ref_")
        L = self.lib
        L.ref_set_opts.argtypes = [C.c_uint, C.c_uint, C.c_uint]
        L.ref_mstab.restype = C.c_uint64
        L.ref_mstab.argtypes = [C.c_ubyte, C.c_uint]
        L.ref_ntread.argtypes = [C.c_char_p, C.c_size_t, _u32p, C.c_uint, _u16p, _u64p]
        L.ref_ntread_batch.restype = C.c_double
        L.ref_ntread_batch.argtypes = [C.c_void_p, _u64p, C.c_size_t, _u32p, C.c_uint, _u16p, _u64p, C.c_int]
        L.ref_compest.argtypes = [_u16p, _dblp, _dblp]
        L.ref_max_threads.restype = C.c_int
        if hasattr(L, "ref_set_gap"):  # an oracle/_ref built before the gap-seed shims lacks them
            L.ref_set_gap.argtypes = [C.c_uint, C.c_uint]
            L.ref_st_hash_seq.restype = C.c_size_t
            L.ref_st_hash_seq.argtypes = [C.c_char_p, C.c_size_t, C.c_uint, _u64p, C.c_size_t]
            L.ref_stread_batch.argtypes = [C.c_void_p, _u64p, C.c_size_t, C.c_uint, _u16p, _u64p, C.c_int]

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def set_opts(self, rBits, sBits, nK):
        self.lib.ref_set_opts(rBits, sBits, nK)

    def mstab(self, c, k):
        return self.lib.ref_mstab(c, k)

    def ntread(self, seq: bytes, kList, sketch, tot):
        kl = np.asarray(kList, dtype=np.uint32)
        self.lib.ref_ntread(seq, len(seq), _ptr(kl, _u32p), len(kl), _ptr(sketch, _u16p), _ptr(tot, _u64p))

    def ntread_batch(self, seqs: np.ndarray, off: np.ndarray, kList, sketch, tot, nthreads=1):
        kl = np.asarray(kList, dtype=np.uint32)
        off = np.ascontiguousarray(off, dtype=np.uint64)
        return self.lib.ref_ntread_batch(seqs.ctypes.data, _ptr(off, _u64p), len(off) - 1, _ptr(kl, _u32p), len(kl),
                                         _ptr(sketch, _u16p), _ptr(tot, _u64p), nthreads)

    def sketch_reads(self, reads, kList, rBits, sBits, nthreads=1):
        self.set_opts(rBits, sBits, len(kList))
        sk = np.zeros(len(kList) * 2 * (1 << rBits), dtype=np.uint16)
        tot = np.zeros(len(kList), dtype=np.uint64)
        lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        buf = np.frombuffer(b"".join(reads) + b"\0", dtype=np.uint8)
        self.ntread_batch(buf, off, kList, sk, tot, nthreads)
        return sk, tot

    def st_hash_seq(self, seq: bytes, k, gap):
        self.lib.ref_set_gap(C.c_uint(k), C.c_uint(gap))
        cap = max(len(seq), 1)
        h = np.zeros(cap, dtype=np.uint64)
        self.lib.ref_st_hash_seq.restype = C.c_size_t
        n = self.lib.ref_st_hash_seq(seq, C.c_size_t(len(seq)), C.c_uint(k), _ptr(h, _u64p), C.c_size_t(cap))
        self.lib.ref_set_gap(C.c_uint(k), C.c_uint(0))
        return h[:n].copy()

    def sketch_reads_gap(self, reads, k, gap, rBits, sBits, nthreads=1):
        """The reference's own stRead with the seed main() builds for -g (ntcard.cpp:407-413)."""
        self.set_opts(rBits, sBits, 1)
        self.lib.ref_set_gap(C.c_uint(k), C.c_uint(gap))
        sk = np.zeros(2 * (1 << rBits), dtype=np.uint16)
        tot = np.zeros(1, dtype=np.uint64)
        lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        buf = np.frombuffer(b"".join(reads) + b"\0", dtype=np.uint8)
        self.lib.ref_stread_batch(buf.ctypes.data, _ptr(off, _u64p), C.c_size_t(len(off) - 1), C.c_uint(k), _ptr(sk, _u16p),
                                  _ptr(tot, _u64p), C.c_int(nthreads))
        self.lib.ref_set_gap(C.c_uint(k), C.c_uint(0))
        return sk, tot

    def compest(self, sketch, rBits, sBits):
        """Full 65535-step recurrence, ~3-4 s (ntcard.cpp:266-272)."""
        self.set_opts(rBits, sBits, 1)
        F0 = C.c_double()
        f = np.zeros(65536, dtype=np.float64)
        self.lib.ref_compest(_ptr(sketch, _u16p), C.byref(F0), _ptr(f, _dblp))
        return F0.value, f

    def max_threads(self):
        return self.lib.ref_max_threads()


class HllReference:
    """The unmodified reference nthll.cpp behind ctypes (oracle/_ref/libnthll_ref.so): its own ntRead / ntComp."""

    def __init__(self, path=HLL_REF_SO):
        self.lib = C.CDLL(path)
        self.lib.ref_hll_set.argtypes = [C.c_uint, C.c_uint]
        self.lib.ref_hll_batch.argtypes = [C.c_void_p, _u64p, C.c_size_t, C.c_void_p, C.c_int]

    @staticmethod
    def available():
        return os.path.exists(HLL_REF_SO)

    def hll_registers(self, reads, k, nBits=16, nthreads=1):
        self.lib.ref_hll_set(k, nBits)
        regs = np.zeros(1 << nBits, dtype=np.uint8)
        lens = np.fromiter((len(r) for r in reads), dtype=np.uint64, count=len(reads))
        off = np.zeros(len(reads) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        buf = np.frombuffer(b"".join(reads) + b"\0", dtype=np.uint8)
        self.lib.ref_hll_batch(buf.ctypes.data, _ptr(off, _u64p), len(off) - 1, regs.ctypes.data, nthreads)
        return regs
