"""scratch: the bench's search leg alone, as bench.py runs it (warm pass over the three profiles, then the timed pass)"""
import json, os, sys
sys.path.insert(0, '.')
import bench
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 1000
per = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ndev = int(sys.argv[3]) if len(sys.argv) > 3 else 1
out = bench.search_leg(list(range(ndev)), per, mbp, 0, False)
print(json.dumps({k: out.get(k) for k in ("value", "seconds", "one_profile_at_a_time", "first_pass_seconds", "hits", "one_gpu", "checks", "contexts_per_gpu")}))
