"""Top source lines by warp-stall samples from an .ncu-rep captured with --import-source on (run where ncu is installed)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
for i, r in enumerate(rows):
    if "Source" in r and any("Sampling" in c for c in r):
        hdr = r; body = rows[i + 1:]; break
if hdr is None:
    print("no source table found; first lines:"); print("\n".join(out.splitlines()[:20])); sys.exit(0)
ci = {c: i for i, c in enumerate(hdr)}
samp = next(c for c in hdr if c.startswith("# Samples") or "Warp Stall Sampling (All" in c)
print("columns:", hdr[:12])
by_line = collections.Counter(); by_op = collections.Counter(); total = 0
for r in body:
    if len(r) != len(hdr): continue
    try: v = float(r[ci[samp]])
    except ValueError: continue
    total += v
    src = r[ci["Source"]].strip()
    by_op[src.split()[0] if src else "?"] += v
    by_line[src[:110]] += v
print("total samples", total)
for k, v in by_line.most_common(40): print(f"{100*v/total:6.2f}%  {k}")
print("--- by opcode / first token")
for k, v in by_op.most_common(25): print(f"{100*v/total:6.2f}%  {k}")
