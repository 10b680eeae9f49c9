"""GPU search == CPU-backend search on a LARGE multi-contig target: the bench's config-4 target generator at <Mbp> (default 100),
three profiles, the GPU side on 8 device contexts (chunks dealt to contexts), the CPU side the oracle's stage calls on all host
threads behind the same host pipeline.  Prints one JSON line: per-profile hit counts, byte-identity of the --tblout tables."""
import hashlib, json, os, sys, time
sys.path.insert(0, '.')
import numpy as np
import bench
from bath_b200 import capi
from oracle import pyoracle as po

mbp = float(sys.argv[1]) if len(sys.argv) > 1 else 100.0
options = {}
for a in sys.argv[2:]:                                  # e.g. std_only=1 block_length=100000
    k, v = a.split("=")
    options[k] = int(v)
models, contigs, plants = bench.search_target(mbp)
ctxs = [capi.Context(0) for _ in range(8)]
t0 = time.perf_counter()
gsecs, gtables, gstats, ghits = bench.run_search(models, contigs, gpu_ctxs=ctxs, **options)
gdt = time.perf_counter() - t0
po.lib(native=True)
be, keep = po.cpu_backend(os.cpu_count() or 1)
t0 = time.perf_counter()
csecs, ctables, cstats, chits = bench.run_search(models, contigs, backends=be, **options)
cdt = time.perf_counter() - t0
same = [g == c for g, c in zip(gtables, ctables)]
cmp = [bench.compare_tables(g, c) for g, c in zip(gtables, ctables)]
# unrounded differences of the hit records, hits matched by target, strand and coordinates (near-equal E-values may swap ranks)
dsc, dbias, dlnp, same_set, swapped = [], [], [], [], []
for gh, ch in zip(ghits, chits):
    key = lambda h: (h["name"], h["ali_from"], h["ali_to"], h["hmm_from"], h["hmm_to"])
    gd, cd = {key(h): h for h in gh}, {key(h): h for h in ch}
    same_set.append(gd.keys() == cd.keys() and len(gd) == len(gh))
    common_keys = gd.keys() & cd.keys()
    dsc.append(max((abs(gd[k]["score"] - cd[k]["score"]) for k in common_keys), default=0.0))
    dbias.append(max((abs(gd[k]["bias"] - cd[k]["bias"]) for k in common_keys), default=0.0))
    dlnp.append(max((abs(gd[k]["lnP"] - cd[k]["lnP"]) for k in common_keys), default=0.0))
    swapped.append(sum(1 for a, b in zip(gh, ch) if key(a) != key(b)))
    worst = sorted(common_keys, key=lambda k: -abs(gd[k]["score"] - cd[k]["score"]))[:3]
    for k in worst:
        print("worst", k, {f: (gd[k][f], cd[k][f]) for f in ("score", "bias", "pre_score", "envsc", "oasc", "lnP")}, gd[k].get("cigar", "")[:60], file=sys.stderr)
os.makedirs("gpurun_out", exist_ok=True)
for k, (g, c) in enumerate(zip(gtables, ctables)):
    if g != c:
        open(f"gpurun_out/parity_gpu_{k}.tbl", "w").write(g)
        open(f"gpurun_out/parity_cpu_{k}.tbl", "w").write(c)
keys = ("pos_past_msv", "pos_past_bias", "pos_past_vit", "pos_past_fwd", "n_orfs", "n_windows", "n_std_windows", "n_regions", "n_multidomain_regions", "n_envelopes", "n_hits_reported")
print(json.dumps({"options": options, "target_mbp": sum(len(d) - 2 for _, d in contigs) / 1e6, "contigs": len(contigs), "profiles": [m.M for m in models],
                  "hits_gpu": [len(h) for h in ghits], "hits_cpu": [len(h) for h in chits], "tables_identical": same, "tables_equivalent": [c[1] for c in cmp],
                  "lines_differing_in_a_last_printed_digit": [c[2] for c in cmp],
                  "same_hit_set": same_set, "hits_at_another_rank": swapped, "max_abs_diff_score_bits": dsc, "max_abs_diff_bias_bits": dbias, "max_abs_diff_lnP": dlnp,
                  "table_sha256": [hashlib.sha256(t.encode()).hexdigest()[:16] for t in gtables],
                  "counters_identical": [all(g[k] == c[k] for k in keys) for g, c in zip(gstats, cstats)],
                  "gpu_seconds": gsecs, "cpu_seconds": csecs, "cpu_threads": os.cpu_count()}))
