"""N > 1 on CPU: two gloo ranks run the cross-rank plumbing of bench.py (bath_b200/ranks.py), and rank 0 runs the sharded search
the way it does under torchrun -- one process, one context per device -- here with two CPU-oracle contexts standing in for two
GPUs; the merged table must equal the one-context table."""
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
    import sys
    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here)); sys.path.insert(0, here)
    import common
    from bath_b200 import hostapi, ranks, synth
    from oracle import pyoracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    t_max = ranks.reduce_scalar(10.0 + rank, "max")
    cells = ranks.reduce_scalar(1000 * (rank + 1), "sum")
    devices = ranks.search_devices(rank, world_size, n_visible=world_size)
    tables = None
    if devices:                                   # rank 0: the whole target over one context per "device"
        model = hostapi.QueryModel(common.golden("tRNA-synthetases.bhmm"), 1)
        contigs, plants = synth.planted_contigs(np.random.default_rng(5), 600_000, [model.mat()], every=20_000, fs_rates=model.fsprob,
                                                min_len=150_000, max_len=300_000)
        tables = []
        for nctx in (1, len(devices)):
            pairs = [pyoracle.cpu_backend(2) for _ in range(nctx)]
            search = hostapi.Search(model, backend=[p[0] for p in pairs], chunk_nt=120_000)
            for name, dsq in contigs:
                search.queue_sequence(name, dsq)
            hits = search.finish()
            tables.append((search.tblout(), search.stats()["nres"], len(hits)))
            search.close()
    dist.barrier()
    q.put((rank, t_max, cells, devices, tables))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reductions_and_rank0_sharded_search():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, t0, c0, d0, tab0), (r1, t1, c1, d1, tab1) = res
    assert t0 == t1 == 11.0                                  # max over ranks
    assert c0 == c1 == 3000.0                                # sum over ranks
    assert d0 == [0, 1] and d1 == [] and tab1 is None       # rank 0 drives both devices, rank 1 none
    one, two = tab0
    assert one[2] >= 20 and one[1] == 2 * 600_000            # hits found; residues = both strands of the whole target
    assert one == two                                         # merged list of two contexts == one context
