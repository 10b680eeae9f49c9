"""scratch: the three profiles of the search leg one after the other (bench.py's way) against the three searches running at once, each
on its own device contexts -- same tables?  faster?"""
import json, sys, threading, time
sys.path.insert(0, '.')
import numpy as np
import bench
from bath_b200 import capi, hostapi

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 1000
ndev = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pers = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "3,4,6,8").split(",")]
models, contigs, plants = bench.search_target(mbp)
pinned = []
for name, dsq in contigs:
    buf = capi.pinned_array(dsq.shape, np.uint8)
    buf[:] = dsq
    pinned.append((name, buf))
total = sum(len(d) - 2 for _, d in contigs)
res = {}
ctxs = [capi.Context(d) for d in range(ndev) for _ in range(8 if ndev <= 2 else max(2, 16 // ndev))]
bench.run_search(models, pinned, gpu_ctxs=ctxs)
secs, tables, _, _ = bench.run_search(models, pinned, gpu_ctxs=ctxs)
res["serial"] = {"contexts": len(ctxs), "seconds": sum(secs), "mbp_s": total * 3 / sum(secs) / 1e6}
for c in ctxs:
    c.close()


def concurrent(per):
    sets = [[capi.Context(d) for d in range(ndev) for _ in range(per)] for _ in models]
    out = [None] * len(models)

    def one(k):
        s = hostapi.Search(models[k], gpu_ctx=sets[k])
        for name, dsq in pinned:
            s.queue_sequence(name, dsq)
        s.finish(fetch=False)
        out[k] = s

    best = None
    for rep in range(3):
        for s in out:
            if s is not None:
                s.close()
        import resource
        r0 = resource.getrusage(resource.RUSAGE_SELF)
        t0 = time.perf_counter()
        th = [threading.Thread(target=one, args=(k,)) for k in range(len(models))]
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        r1 = resource.getrusage(resource.RUSAGE_SELF)
        cpu = (r1.ru_utime - r0.ru_utime) + (r1.ru_stime - r0.ru_stime)
        if rep >= 1 and (best is None or dt < best):
            best = dt; best_cpu = cpu
    tb = [s.tblout(header=False) for s in out]
    for s in out:
        s.close()
    for cs in sets:
        for c in cs:
            c.close()
    return {"contexts_per_profile_per_gpu": per, "seconds": best, "cpu_seconds": best_cpu, "mbp_s": total * 3 / best / 1e6, "tables_identical_to_serial": tb == tables}


for per in pers:
    res[f"concurrent_{per}"] = concurrent(per)
print(json.dumps(res))
