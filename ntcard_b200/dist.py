"""Multi-GPU plumbing of the sketch path: one process per GPU, reads sharded over ranks, ONE sum
all-reduce of the uint32 device counters (and of the per-k totals) at the end.

The reference has no distributed mode; its only reduction is `totalKmers[k] += totKmer[k]`
(ntcard.cpp:464-466) on a sketch shared by all threads.  Here each rank owns a full-size sketch; the
sum is order independent, and because device counters are uint32 and narrowed mod 2^16 only after
the reduction, the result equals the reference's wrapping uint16 counters exactly:
(sum_r c_r mod 2^32) mod 2^16 == (sum_r c_r) mod 2^16.

NCCL has no 16-bit integer type, hence the uint32 wire format (int32 on the torch side: two's
complement addition is the same operation).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous shard [lo, hi) of n_total reads for `rank` (read i lives on GPU i // ceil(n/world))."""
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def all_reduce_sketch(counters: torch.Tensor, totals) -> np.ndarray:
    """In-place sum all-reduce of the counters tensor (int32 view of uint32 counters, on the GPU with
    NCCL or on the CPU with gloo) and of the per-k k-mer totals.  Returns the reduced totals (uint64)."""
    tot = torch.as_tensor(np.asarray(totals, dtype=np.uint64).astype(np.int64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
        tot = tot.to(counters.device)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    return tot.cpu().numpy().astype(np.uint64)


def narrow_counters(counters_u32: np.ndarray) -> np.ndarray:
    """uint32 -> uint16 mod 2^16: the reference's `uint16_t ++` wrap (ntcard.cpp:133,143)."""
    return (counters_u32 & 0xFFFF).astype(np.uint16)


def reduce_scatter_hist(sketch, counters: torch.Tensor, rBits: int):
    """The cheaper reduction when only F0 / f_i are wanted: compEst needs the counter-VALUE histogram, not
    the table.  reduce-scatter the uint32 counters (each rank receives the summed 1/N slice: half the wire
    traffic of an all-reduce), histogram the narrowed slice on the device, all-reduce the 512 KiB/k
    histograms.  Returns p_hist uint32 [nK, 2, 65536] (identical on all ranks), ready for estimate()."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    n = counters.numel()
    if world == 1:
        p = sketch.hist_range(counters.data_ptr(), 0, n)
    else:
        rank = dist.get_rank()
        assert n % world == 0
        mine = torch.empty(n // world, dtype=counters.dtype, device=counters.device)
        dist.reduce_scatter_tensor(mine, counters, op=dist.ReduceOp.SUM)
        torch.cuda.current_stream(counters.device).synchronize()   # the histogram kernel runs on the sketch's own stream
        p = sketch.hist_range(mine.data_ptr(), rank * (n // world), n // world)
        pt = torch.from_numpy(p.astype(np.int64)).to(counters.device)
        dist.all_reduce(pt, op=dist.ReduceOp.SUM)
        p = pt.cpu().numpy().astype(np.uint32)
    p[:, :, 0] = (1 << rBits) - p[:, :, 1:].sum(axis=2, dtype=np.uint64).astype(np.uint32)
    return p


# ---- sparse reduction: exchange the hit log, not the sketch ------------------------------------------------
def plan_exchange(counts: np.ndarray, rank: int):
    """counts[q, s] = log blocks rank q holds for slice s.  Slice s is owned by rank s % world.  Returns
    (owned uint8[n_slices], send_slices, send_splits[world], recv_splits[world], runs[(slice, n_blocks)]):
    this rank sends, to every other rank d, the blocks of d's slices in increasing slice order, and receives
    from every other rank q the blocks of its own slices in the same order (runs, in arrival order)."""
    world, n_slices = counts.shape
    owner = np.arange(n_slices) % world
    owned = (owner == rank).astype(np.uint8)
    mine = counts[rank]
    send_slices, send_splits, recv_splits, runs = [], [], [], []
    for d in range(world):
        sl = np.nonzero((owner == d) & (mine > 0))[0] if d != rank else np.zeros(0, dtype=np.int64)
        send_slices += [int(x) for x in sl]
        send_splits.append(int(mine[sl].sum()))
    own = np.nonzero(owner == rank)[0]
    for q in range(world):
        tot = 0
        if q != rank:
            sl = own[counts[q, own] > 0]
            runs += [(int(s), int(counts[q, s])) for s in sl]
            tot = int(counts[q, sl].sum())
        recv_splits.append(tot)
    return owned, send_slices, send_splits, recv_splits, runs


def exchange_feasible(counts: np.ndarray, info: np.ndarray) -> bool:
    """Every rank evaluates the same predicate on the same gathered data: do the blocks every rank is about to
    receive fit its pool and its per-slice block lists?  info[q] = (blocks used, blocks in pool, list capacity)."""
    world, n_slices = counts.shape
    if n_slices < world:
        return False
    for r in range(world):
        incoming = sum(int(counts[q, s]) for q in range(world) if q != r for s in range(r, n_slices, world))
        if int(info[r, 0]) + incoming > int(info[r, 1]):
            return False
        for s in range(r, n_slices, world):
            if int(counts[:, s].sum()) > int(info[r, 2]):
                return False
    return True


def exchange_hist(sketch, rBits: int, device, totals_out=None):
    """The cheap reduction for inputs whose hit log is small against the sketch (10 M reads: 74 MB vs 1 GiB per
    k): ranks swap log blocks so that rank r holds everything sampled into the slices it owns (one all-to-all),
    each rank zeroes + applies + histograms ONLY its slices, and the 512 KiB/k histograms are all-reduced.
    Returns p_hist uint32 [nK, 2, 65536] (identical on all ranks), or None when some rank's log is no longer
    complete (it was flushed or overflowed) or the blocks would not fit -- then use reduce_scatter_hist.
    totals_out: optional list; receives the per-k F1 summed over ranks (it rides on the same all-gather)."""
    import os
    import time
    prof = os.environ.get("NTC_DIST_PROFILE") and dist.get_rank() == 0
    t = [time.perf_counter()]

    def lap(name):
        if prof:
            torch.cuda.synchronize(device)
            t.append(time.perf_counter())
            print(f"  exchange_hist {name}: {1e3 * (t[-1] - t[-2]):.3f} ms", flush=True)

    world, rank = dist.get_world_size(), dist.get_rank()
    nblk, ok, info = sketch.log_counts()
    f1 = sketch.totals_nosync()                            # final: log_counts waited for the scan kernels
    lap("log_counts (sync)")
    n_slices = len(nblk)
    mine = torch.from_numpy(np.concatenate([nblk.astype(np.int64), np.asarray(info, dtype=np.int64), [1 if ok else 0],
                                            f1.astype(np.int64)])).to(device)
    allv = torch.empty(world * mine.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allv, mine)
    allv = allv.cpu().numpy().reshape(world, -1)
    lap("all_gather counts")
    counts, infos, oks = allv[:, :n_slices], allv[:, n_slices:n_slices + 3], allv[:, n_slices + 3]
    if totals_out is not None:
        totals_out.append(allv[:, n_slices + 4:].sum(axis=0).astype(np.uint64))
    if not oks.all() or not exchange_feasible(counts, infos):
        return None
    owned, send_slices, send_splits, recv_splits, runs = plan_exchange(counts, rank)
    W = 260                                                # words per block on the wire (256 entries + fill + pad)
    n_send, n_recv = sum(send_splits), sum(recv_splits)
    send = _buffer(device, "send", n_send * W)
    recv = _buffer(device, "recv", n_recv * W)
    hist = _buffer(device, "hist", sketch.nK * 2 * 65536)
    lap("plan + buffers")
    same_stream = getattr(sketch, "stream_handle", None) == torch.cuda.current_stream(device).cuda_stream
    sketch.log_export(send_slices, send.data_ptr())
    if not same_stream:
        sketch.stream_sync()                               # the export ran on the sketch's own stream
    lap("export")
    dist.all_to_all_single(recv[:n_recv * W], send[:n_send * W], [x * W for x in recv_splits], [x * W for x in send_splits])
    if not same_stream:
        torch.cuda.current_stream(device).synchronize()    # the import runs on the sketch's own stream
    lap(f"all_to_all ({n_send * W * 4 / 1e6:.1f} MB out)")
    sketch.log_import(recv.data_ptr(), n_recv, runs)
    sketch.flush_slices(owned)
    sketch.hist_slices(owned, d_out=hist.data_ptr())
    if not same_stream:
        sketch.stream_sync()
    lap("import + flush_slices + hist_slices")
    h = hist[:sketch.nK * 2 * 65536]
    dist.all_reduce(h, op=dist.ReduceOp.SUM)               # int32 on the device: a bin counts at most 2^rBits buckets
    p = h.cpu().numpy().view(np.uint32).reshape(sketch.nK, 2, 65536).copy()
    p[:, :, 0] = (1 << rBits) - p[:, :, 1:].sum(axis=2, dtype=np.uint64).astype(np.uint32)
    lap("all_reduce hist")
    return p


# ---- reduction over peer memory: no host planning, no host synchronisation -----------------------------------------
def peer_setup(sketch):
    """Once per context: all-gather the IPC handles of the ranks' hit logs and map them (ntc_peer_attach).  All ranks must
    be processes on ONE node (CUDA IPC + NVLink peer access)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = sketch.peer_export()
    gathered = [None] * world
    dist.all_gather_object(gathered, mine.tobytes())
    sketch.peer_attach(world, rank, np.frombuffer(b"".join(gathered), dtype=np.uint8))


class PeerReducer:
    """The one reduction at the end of a multi-GPU run (the reference has a single shared sketch, ntcard.cpp:439, and an F1
    merge, ntcard.cpp:464-466), host-free: two small NCCL all-reduces around ONE kernel sequence that pulls the other
    ranks' hit-log blocks through NVLink peer memory (ntc_reduce_owned):

        status := [F1 per k, my log is incomplete]      (kernel)          all-reduce(status)   -- F1 totals + barrier
        owned slices := zeros + every rank's entries     (apply_owned_kernel, peer loads)
        hist := counter-value histogram of my slices    (kernel)          all-reduce(hist)     -- result + barrier

    Everything is ordered on torch's current stream, which must be the sketch's stream.  reduce() returns device tensors;
    result() is the only host synchronisation (and checks the fallback flag)."""

    def __init__(self, sketch, device):
        self.sk, self.device = sketch, device
        self.nK = sketch.nK
        self.status = torch.zeros(self.nK + 1, dtype=torch.int64, device=device)
        self.hist = torch.zeros(self.nK * 2 * 65536, dtype=torch.int32, device=device)
        # every rank must be able to map every other rank's log (CUDA IPC + peer access); all ranks take the same decision
        try:
            peer_setup(sketch)
            mine, self.error = 1, None
        except Exception as e:  # e.g. no peer access between two devices
            mine, self.error = 0, e
        flag = torch.tensor([mine], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.ok = bool(flag.item())

    def reduce(self):
        self.sk.log_status_device(self.status.data_ptr())
        dist.all_reduce(self.status, op=dist.ReduceOp.SUM)
        self.sk.reduce_owned(self.status.data_ptr(), self.hist.data_ptr())
        dist.all_reduce(self.hist, op=dist.ReduceOp.SUM)
        return self.status, self.hist

    def result(self, rBits):
        """(p_hist uint32 [nK, 2, 65536] or None when some rank's log was incomplete -> dense fallback, F1 uint64 [nK])"""
        st = self.status.cpu().numpy()
        f1 = st[:self.nK].astype(np.uint64)
        if st[self.nK] != 0:
            return None, f1
        p = self.hist.cpu().numpy().view(np.uint32).reshape(self.nK, 2, 65536).copy()
        p[:, :, 0] = (1 << rBits) - p[:, :, 1:].sum(axis=2, dtype=np.uint64).astype(np.uint32)
        return p, f1


_BUFFERS = {}


def _buffer(device, name, n_words):
    """Reusable int32 device buffers of the exchange (grown with slack, never shrunk)."""
    key = (str(device), name)
    buf = _BUFFERS.get(key)
    if buf is None or buf.numel() < max(n_words, 1):
        buf = torch.empty(max(int(n_words * 1.25), 1024), dtype=torch.int32, device=device)
        _BUFFERS[key] = buf
    return buf


def all_reduce_hll(registers: torch.Tensor) -> torch.Tensor:
    """nthll's only reduction (nthll.cpp:234-239: tVec[j] = max over threads of mVec[j]), across ranks: an in-place MAX
    all-reduce of the uint8 HyperLogLog registers (the tensor HllSketch was created on, or a CPU tensor with gloo).
    64 KiB at the default 2^16 registers: one latency-bound collective; NCCL and gloo both reduce uint8."""
    if registers.dtype != torch.uint8:
        raise TypeError("HyperLogLog registers are uint8")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(registers, op=dist.ReduceOp.MAX)
    return registers
