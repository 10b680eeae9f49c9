"""scratch: symbolise the samples of prof_preload.c (self time per function, per shared object)"""
import bisect, collections, subprocess, sys
maps, samples = [], []
for line in open(sys.argv[1]):
    if line.startswith("M "):
        f = line[2:].split()
        lo, hi = [int(x, 16) for x in f[0].split("-")]
        maps.append((lo, hi, int(f[2], 16), f[5] if len(f) > 5 else "[anon]"))
    elif line.startswith("S "):
        samples.append(int(line[2:], 16))
maps.sort()
syms = {}
def table(path):
    if path not in syms:
        out = []
        try:
            for l in subprocess.run(["nm", "-n", "-C", "--defined-only", path], capture_output=True, text=True).stdout.splitlines():
                p = l.split(None, 2)
                if len(p) == 3 and p[1] in "tTwW":
                    out.append((int(p[0], 16), p[2]))
        except Exception:
            pass
        syms[path] = out
    return syms[path]
per_obj, per_fn = collections.Counter(), collections.Counter()
for a in samples:
    i = bisect.bisect_right(maps, (a, 1 << 62, 0, "")) - 1
    if i < 0 or not (maps[i][0] <= a < maps[i][1]):
        per_obj["?"] += 1
        continue
    lo, hi, off, path = maps[i]
    per_obj[path.split("/")[-1]] += 1
    want = [w for w in sys.argv[2:]] or ["libbathhost"]
    if any(w in path for w in want):
        t = table(path)
        j = bisect.bisect_right(t, (a - lo + off, "\xff")) - 1
        per_fn[(path.split("/")[-1], t[j][1][:110] if j >= 0 else "?")] += 1
n = len(samples)
print(f"{n} samples (1 ms of process CPU each)")
for k, v in per_obj.most_common(12):
    print(f"  {100 * v / n:5.1f} %  {k}")
print("self time by function:")
for (o, fn), v in per_fn.most_common(40):
    print(f"  {100 * v / n:5.1f} %  {v:6d}  {o}  {fn}")
