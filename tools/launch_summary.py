"""ncu launch list (--metrics gpu__time_duration.sum --csv) of tools/train_once.py -> per-kernel summary of ONE training step
(the launches between the last two clip_adam_kernel launches).   python tools/launch_summary.py <csv> [title]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
data = [(r[ki], float(r[vi].replace(",", "")), r[gi]) for r in rows[hdr + 1:] if len(r) > vi]
idx = [i for i, (n, _, _) in enumerate(data) if "clip_adam_kernel" in n]
step = data[idx[-2] + 1: idx[-1] + 1]
short = lambda n: n.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[-60:]
tot = sum(d[1] for d in step)
print(sys.argv[2] if len(sys.argv) > 2 else "")
print(f"one training step: {len(step)} launches, {tot / 1000:.1f} us serialised under ncu (cold-cache, one kernel at a time: shares matter)\n")
agg = collections.OrderedDict()
for n, v, g in step:
    a = agg.setdefault(short(n), [0.0, 0])
    a[0] += v
    a[1] += 1
for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v / 1000:9.1f} us {c:3d}x {100 * v / tot:5.1f}%  {n}")
print("\nin launch order:")
for n, v, g in step:
    print(f"{v / 1000:9.1f} us  grid {g:>14}  {short(n)}")
