"""Per-opcode / per-instruction stall summary of an `ncu --page source --csv` dump (one kernel)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
H = rows[1]
data = [r for r in rows[2:] if len(r) == len(H) and r[H.index('# Samples')].isdigit()]
si, src, ie = H.index('# Samples'), H.index('Source'), H.index('Instructions Executed')
cols = ['stall_long_sb', 'stall_not_selected', 'stall_selected', 'stall_wait', 'stall_dispatch', 'stall_math',
        'stall_short_sb', 'stall_mio', 'stall_lg', 'stall_barrier', 'stall_no_inst', 'stall_branch_resolving']
ci = {c: H.index(c) for c in cols}
tot = sum(int(r[si]) for r in data)
print("total samples", tot, "SASS instructions", len(data))
print({c: sum(int(r[ci[c]]) for r in data) for c in cols})
op, exe = collections.Counter(), collections.Counter()
for r in data:
    t = r[src].split()
    o = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    op[o] += int(r[si])
    exe[o] += int(r[ie])
print("opcode: samples, warp-instructions executed")
for o, c in op.most_common(14):
    print(f"  {o:10s} {c:8d} {exe[o]:12d}")
print("top instructions by samples (samples, long_sb, wait, source)")
for r in sorted(data, key=lambda r: -int(r[si]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"  {r[si]:>6s} {r[ci['stall_long_sb']]:>6s} {r[ci['stall_wait']]:>6s}  {r[src][:100]}")
