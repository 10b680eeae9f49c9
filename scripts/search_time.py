"""scratch: stage timing of the host pipeline on a synthetic genome"""
import sys, time, numpy as np
sys.path.insert(0, '.')
from bath_b200 import capi, hostapi, synth
mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 20
m = hostapi.QueryModel('tests/golden/tRNA-synthetases.bhmm', 1)
rng = np.random.default_rng(42)
d, plants = synth.planted_genome(rng, int(mbp * 1e6), m.mat(), every=50000, fs_rate=m.fsprob)
ctx = capi.Context(0)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    s = hostapi.Search(m, ctx)
    t = time.perf_counter(); s.add_sequence('g', d); hits = s.finish(); dt = time.perf_counter() - t
    st = s.stats()
    print(f"{mbp} Mbp in {dt:.3f} s = {mbp/dt:.1f} Mbp/s, hits {len(hits)}")
    print({k: v for k, v in st.items() if k.startswith('us_')})
    print({k: v for k, v in st.items() if not k.startswith('us_')})
