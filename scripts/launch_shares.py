"""Per-kernel share of the device time in an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/*_launches.csv)."""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).strip()
    ns = float(r[-1].replace(",", ""))
    unit = r[-2]
    ms = ns / 1e6 if unit in ("ns", "nsecond") else ns / 1e3 if unit in ("us", "usecond") else ns if unit in ("ms", "msecond") else ns * 1e3
    tot[name] += ms; cnt[name] += 1
s = sum(tot.values())
print(f"# kernel launches of `{' '.join(sys.argv[2:])}` under ncu --metrics gpu__time_duration.sum")
print("# (cold-cache, serialised: shares, not absolute times)")
for k, v in sorted(tot.items(), key=lambda t: -t[1]):
    print(f"{100 * v / s:6.2f}%  {v:10.3f} ms  x{cnt[k]:<4d} {k}")
