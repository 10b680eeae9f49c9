"""scratch: one profile after the other over <per> contexts on each of <ndev> devices (config 4's target), BATHHOST_TRACE phase times
of the LAST timed profile pass on stderr"""
import json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
import bench
from bath_b200 import capi
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 1000
per = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ndev = int(sys.argv[3]) if len(sys.argv) > 3 else 1
models, contigs, plants = bench.search_target(mbp)
pinned = []
for name, dsq in contigs:
    buf = capi.pinned_array(dsq.shape, np.uint8)
    buf[:] = dsq
    pinned.append((name, buf))
ctxs = [capi.Context(d) for d in range(ndev) for _ in range(per)]
bench.run_search(models, pinned, gpu_ctxs=ctxs)
print("==== timed pass ====", file=sys.stderr, flush=True)
secs, tables, stats, hits = bench.run_search(models, pinned, gpu_ctxs=ctxs)
print(json.dumps({"seconds_per_profile": secs, "contexts": len(ctxs)}))
