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


# ---- sparse reduction (hit-log exchange): host planning --------------------------------------------------
def test_plan_exchange_is_consistent_across_ranks():
    """What rank q plans to send to rank r must be exactly what r plans to receive from q, slice by slice."""
    from ntcard_b200.dist import exchange_feasible, plan_exchange
    rng = np.random.default_rng(3)
    for world, n_slices in ((2, 32), (4, 64), (8, 32), (3, 7)):
        counts = rng.integers(0, 5, size=(world, n_slices)).astype(np.int64)
        counts[rng.random(counts.shape) < 0.3] = 0
        plans = [plan_exchange(counts, r) for r in range(world)]
        for r in range(world):
            owned, send_slices, send_splits, recv_splits, runs = plans[r]
            assert [int(x) for x in owned] == [int(s % world == r) for s in range(n_slices)]
            assert send_splits[r] == 0 and recv_splits[r] == 0
            assert sum(send_splits) == sum(int(counts[r, s]) for s in send_slices)
            assert sum(recv_splits) == sum(n for _, n in runs)
            # the runs rank r expects, grouped by source rank, are the sources' send lists restricted to r's slices
            pos = 0
            for q in range(world):
                if q == r:
                    continue
                q_sends_to_r = [(s, int(counts[q, s])) for s in plans[q][1] if s % world == r]
                got = []
                tot = 0
                while tot < recv_splits[q]:
                    got.append(runs[pos])
                    tot += runs[pos][1]
                    pos += 1
                assert got == q_sends_to_r and tot == plans[q][2][r]
            assert pos == len(runs)
        info = np.tile(np.array([[10, 10_000, 1_000]]), (world, 1))
        assert exchange_feasible(counts, info) == (n_slices >= world)
        info[0, 1] = 11                                      # rank 0's pool is (nearly) full
        incoming0 = sum(int(counts[q, s]) for q in range(1, world) for s in range(0, n_slices, world))
        assert exchange_feasible(counts, info) == (n_slices >= world and incoming0 <= 1)


def _exchange_worker(rank, world, port, out_dir):
    """all_to_all of fake log blocks over gloo: every rank ends up with exactly the blocks of its own slices."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ntcard_b200.dist import plan_exchange
    n_slices, E = 8, 4
    rng = np.random.default_rng(11)
    counts = rng.integers(0, 4, size=(world, n_slices)).astype(np.int64)          # same on every rank
    # block b of slice s on rank q carries the values 1000 q + 10 s + b (E copies)
    blocks = {s: [np.full(E, 1000 * rank + 10 * s + b, dtype=np.int32) for b in range(counts[rank, s])] for s in range(n_slices)}
    owned, send_slices, send_splits, recv_splits, runs = plan_exchange(counts, rank)
    send = np.concatenate([blk for s in send_slices for blk in blocks[s]] + [np.zeros(0, dtype=np.int32)])
    recv = torch.zeros(sum(recv_splits) * E, dtype=torch.int32)
    dist.all_to_all_single(recv, torch.from_numpy(send.copy()), [x * E for x in recv_splits], [x * E for x in send_splits])
    recv = recv.numpy()
    ok, pos = True, 0
    srcs = [q for q in range(world) if q != rank]
    per_src = {q: [(s, int(counts[q, s])) for s in range(n_slices) if s % world == rank and counts[q, s] > 0] for q in srcs}
    flat = [(q, s, n) for q in srcs for s, n in per_src[q]]
    ok &= [(s, n) for _, s, n in flat] == runs
    for q, s, n in flat:
        for b in range(n):
            ok &= bool((recv[pos * E:(pos + 1) * E] == 1000 * q + 10 * s + b).all())
            pos += 1
    ok &= pos * E == len(recv)
    with open(os.path.join(out_dir, f"x{rank}.txt"), "w") as f:
        f.write(f"{int(ok)}\n")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_block_exchange(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_exchange_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(open(os.path.join(str(tmp_path), f"x{r}.txt")).read().strip() == "1" for r in range(world))
