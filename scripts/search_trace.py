"""scratch: the bench's search leg alone (config 4: 3 profiles x 1 Gbp), with BATHHOST_TRACE phase times on stderr"""
import json, os, sys
sys.path.insert(0, '.')
os.environ.setdefault("BATHHOST_TRACE", "1")
import bench
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 1000
per = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ndev = int(sys.argv[3]) if len(sys.argv) > 3 else 1
out = bench.search_leg(list(range(ndev)), per, mbp, 0, False)
print(json.dumps({k: out.get(k) for k in ("value", "seconds_per_profile", "first_pass_seconds", "hits", "stats_per_profile", "one_gpu", "checks")}))
