"""N > 1 path on CPU: two gloo ranks exercise the sharding and the cross-rank bookkeeping (no collective on the DP path)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, q):
    import torch.distributed as dist
    from bath_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    blocks = shard.shard_blocks(11, rank, world_size)
    t_max = shard.reduce_scalar(10.0 + rank, "max")
    nres = shard.reduce_scalar(1000 * (rank + 1), "sum")
    hits = [{"name": f"seq{rank}", "strand": 1, "ali_from": 5 + rank, "ali_to": 90, "lnP_raw": -30.0 - rank},
            {"name": f"seq{rank}", "strand": -1, "ali_from": 500, "ali_to": 400, "lnP_raw": -1.0}]
    gathered = shard.gather_hits(hits)
    merged = shard.merge_hits(gathered, nres, 167) if rank == 0 else None
    q.put((rank, blocks, t_max, nres, merged))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, t0, n0, m0), (r1, b1, t1, n1, m1) = res
    assert sorted(b0 + b1) == list(range(11)) and not set(b0) & set(b1)      # every block searched exactly once
    assert t0 == t1 == 11.0                                                   # max over ranks
    assert n0 == n1 == 3000.0                                                 # residue count = E-value search space
    assert m1 is None and [h["name"] for h in m0] == ["seq1", "seq0", "seq0", "seq1"][: len(m0)] or len(m0) >= 2
    # lnP = lnP_raw + log(N / W), W = 3 * max_length (src/p7_tophits.c:795)
    assert abs(m0[0]["lnP"] - (-31.0 + float(np.log(np.float32(3000) / np.float32(501))))) < 1e-6
    assert all(m0[i]["lnP"] <= m0[i + 1]["lnP"] for i in range(len(m0) - 1))
