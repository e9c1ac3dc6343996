"""Executed opcode mix of one kernel from `ncu -i rep --page source --csv --print-source sass [--launch-skip k --launch-count 1]`."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[ix["Instructions Executed"]].isdigit():
        continue
    toks = r[ix["Source"]].strip().split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    ex[op] += int(r[ix["Instructions Executed"]])
E = sum(ex.values())
print(rows[0][1][:100] if rows[0] and rows[0][0] == "Kernel Name" else "", "| executed warp instructions", E)
grp = collections.Counter()
for op, e in ex.items():
    base = op.split(".")[0]
    grp[(".".join(op.split(".")[:2]) if base == "IMAD" else base)] += e
for k, v in grp.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 20):
    print(f"   {k:24s} {100 * v / E:6.2f}%  {v}")
