"""world_size-2 gloo test (CPU) of the multi-GPU host logic: shard the reads, build one sketch per
rank, sum all-reduce the uint32 counters and the totals, narrow mod 2^16 -- the result must equal
the single-process sketch of all reads, including counters that wrap past 65535."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ntcard_b200.dist import all_reduce_sketch, narrow_counters, shard_range
    from oracle.pyoracle import Oracle
    orc = Oracle()
    kList, rBits, sBits = [12, 32], 12, 1
    wrap_read = orc.gen_read(9, 0, 40, 0, 0)
    a = orc.gen_reads(5, 0, 3000, 100, 1, 300)
    reads = [bytes(a[i * 100:(i + 1) * 100]) for i in range(3000)] + [wrap_read] * 70000   # wraps a uint16 counter
    lo, hi = shard_range(len(reads), rank, world)
    # per-rank sketch: the checker stands in for the device here (host plumbing is what is under test);
    # widen to uint32 the way the device holds it, un-wrapping nothing: each shard alone stays below 65536
    sk16, tot = orc.sketch_reads(reads[lo:hi], kList, rBits, sBits)
    assert int(sk16.max()) < 65535
    ctr = torch.from_numpy(sk16.astype(np.uint32).view(np.int32).copy())
    tot_all = all_reduce_sketch(ctr, tot)
    got = narrow_counters(ctr.numpy().view(np.uint32))
    want, want_tot = orc.sketch_reads(reads, kList, rBits, sBits)
    ok = np.array_equal(got, want) and np.array_equal(tot_all, want_tot)
    wrapped = bool((ctr.numpy().view(np.uint32) > 65535).any())
    with open(os.path.join(out_dir, f"r{rank}.txt"), "w") as f:
        f.write(f"{int(ok)} {int(wrapped)} {lo} {hi}\n")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduce(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rows = [open(os.path.join(str(tmp_path), f"r{r}.txt")).read().split() for r in range(world)]
    assert all(r[0] == "1" for r in rows), rows
    assert all(r[1] == "1" for r in rows), "the test input must exercise the mod-2^16 wrap"
    assert rows[0][2] == "0" and rows[0][3] == rows[1][2]


def test_shard_range_covers_everything():
    from ntcard_b200.dist import shard_range
    for n in (0, 1, 7, 8, 1000, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
