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
        p = sketch.hist_range(mine.data_ptr(), rank * (n // world), n // world)
        pt = torch.from_numpy(p.astype(np.int64)).to(counters.device)
        dist.all_reduce(pt, op=dist.ReduceOp.SUM)
        p = pt.cpu().numpy().astype(np.uint32)
    p[:, :, 0] = (1 << rBits) - p[:, :, 1:].sum(axis=2, dtype=np.uint64).astype(np.uint32)
    return p
