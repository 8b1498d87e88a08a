#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on N GPUs): shard reads over ranks, sketch each shard on its GPU,
then compare BOTH reductions against the oracle's sketch of all reads:
  (a) all-reduce of the uint32 counters + finish (full table, bit-exact),
  (b) reduce-scatter + per-rank histogram + histogram all-reduce (the dense fallback),
  (c) the reduction over NVLink peer memory (ntcard_b200.dist.PeerReducer: every rank applies all ranks' hit-log entries of the
      slices it owns, read through peer pointers; what bench.py times).
Prints one line per rank 0.   torchrun --nproc-per-node 2 tools/check_multi_gpu.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ntcard_b200 as nt  # noqa: E402
from ntcard_b200.dist import PeerReducer, all_reduce_sketch, reduce_scatter_hist, shard_range  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    orc = Oracle()
    kList, rBits, sBits, L, n = [32, 64], 25, 7, 150, 400_000   # 16 sketch slices: enough for 8 ranks
    wrap = orc.gen_read(9, 0, 40, 0, 0)
    a = orc.gen_reads(3, 0, n, L, 1, n // 8)
    lo, hi = shard_range(n, rank, world)
    stride = nt.stride_words(L)
    counters = torch.zeros(len(kList) * 2 << rBits, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st), nt.Sketch(kList, rBits=rBits, sBits=sBits, device=lr, d_counters=counters.data_ptr(), stream=st.cuda_stream) as sk:
        shard = nt.gen_packed(3, lo, hi - lo, L, 1, n // 8, stride)
        red = PeerReducer(sk, dev)
        for _ in range(2):                                                       # twice: reset + reuse
            sk.reset()
            sk.submit(shard, None, hi - lo, stride)
            red.reduce()                                                         # (c) peer memory
            p_ex, f1_ex = red.result(rBits)
        assert p_ex is not None, "the hit log should be complete here"
        sk.reset()
        sk.submit(shard, None, hi - lo, stride)
        sk.sync()
        p_rs = reduce_scatter_hist(sk, counters.clone(), rBits)                 # (b)
        tot = all_reduce_sketch(counters, sk.totals())                           # (a)
        sk.set_totals(tot)
        t, f1, p_ar = sk.finish(counters=True, hist=True)
    if rank == 0:
        reads = [bytes(a[i * L:(i + 1) * L]) for i in range(n)]
        want, wf1 = orc.sketch_reads(reads, kList, rBits, sBits, nthreads=8)
        ok_a = bool(np.array_equal(t.reshape(-1), want) and np.array_equal(f1, wf1))
        ok_b = bool(np.array_equal(p_rs, p_ar))
        ok_x = bool(np.array_equal(p_ex, p_ar) and np.array_equal(f1_ex, wf1))
        est = [nt.estimate(p_hist=p_rs[ki], rBits=rBits, sBits=sBits, covMax=16) for ki in range(len(kList))]
        oest = [orc.compest(np.ascontiguousarray(want[ki * (2 << rBits):(ki + 1) * (2 << rBits)]), None, rBits, sBits, 16) for ki in range(len(kList))]
        ok_c = all(e[0] == o[0] and np.array_equal(e[1][1:], o[1][1:17]) for e, o in zip(est, oest))
        print(f"multi-gpu check world={world}: allreduce sketch bit-exact={ok_a} reduce-scatter hist == allreduce hist={ok_b} peer-memory reduction hist == allreduce hist={ok_x} "
              f"F0/f_i equal oracle={ok_c} F1={[int(x) for x in f1]} F0={[e[0] for e in est]}", flush=True)
        assert ok_a and ok_b and ok_c and ok_x
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
