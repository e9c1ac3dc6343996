"""Aggregate an `ncu --page source --csv --print-source sass` dump: samples by opcode class and the top stalled instructions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); stalls = collections.Counter(); ex = collections.Counter()
recs = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    n = int(r[ix["# Samples"]] or 0)
    e = int(r[ix["Instructions Executed"]] or 0)
    tot[op] += n; ex[op] += e
    for k in ("stall_long_sb", "stall_wait", "stall_math", "stall_no_inst", "stall_short_sb", "stall_dispatch", "stall_selected", "stall_not_selected", "stall_branch_resolving", "stall_lg", "stall_mio"):
        stalls[k] += int(r[ix[k]] or 0)
    recs.append((n, e, r[ix["Address"]], src, r[ix["stall_long_sb"]], r[ix["stall_wait"]], r[ix["stall_math"]]))
S = sum(tot.values()); E = sum(ex.values())
print("total samples", S, "instructions executed", E)
print("stall mix:", {k: round(v / S, 3) for k, v in stalls.most_common()})
print("by opcode: samples%  executed%")
for op, n in tot.most_common(18):
    print(f"  {op:22s} {100*n/S:6.2f}  {100*ex[op]/E:6.2f}")
print("top instructions:")
for n, e, a, src, lsb, w, m in sorted(recs, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"  {n:6d} ex={e:9d} {a[-6:]} long_sb={lsb:>5s} wait={w:>5s} math={m:>5s}  {src[:90]}")
