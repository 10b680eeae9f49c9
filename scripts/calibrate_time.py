"""Times the calibration of the 12 models of tRNA-proteins.bhmm (bathconvert flow + the full p7_Calibrate flow) through the device
library against the same host code over the CPU oracle's stage calls.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bath_b200 import capi, hostapi          # noqa: E402
from oracle import pyoracle                  # noqa: E402  (CPU comparison leg only)

path = os.path.join(ROOT, "tests", "golden", "tRNA-proteins.bhmm")
models = [hostapi.QueryModel(path, i) for i in range(12)]
ctx = capi.Context(0)


def run(**kw):
    out = []
    for flow in (True, False):
        x, t0, evs = 0, time.perf_counter(), []
        for m in models:
            ev, x = hostapi.calibrate(m, convert_flow=flow, rng_state=x if flow else 0, **kw)
            evs.append(ev)
        out.append((time.perf_counter() - t0, evs))
    return out


run(gpu_ctx=ctx)                             # warm-up: module load, allocations
gpu = run(gpu_ctx=ctx)
be, keep = pyoracle.cpu_backend(os.cpu_count() or 1)
cpu = run(backend=be)
diff = max(abs(a - b) for (_, ge), (_, ce) in zip(gpu, cpu) for g, c in zip(ge, ce) for a, b in zip(g, c))
print(json.dumps({"models": 12, "M": [m.M for m in models],
                  "gpu_seconds": {"bathconvert_flow_fs3_fs5": gpu[0][0], "p7_Calibrate_flow_all5": gpu[1][0]},
                  "cpu_oracle_seconds": {"bathconvert_flow_fs3_fs5": cpu[0][0], "p7_Calibrate_flow_all5": cpu[1][0], "threads": os.cpu_count()},
                  "max_abs_diff_gpu_vs_cpu_oracle": diff}))
