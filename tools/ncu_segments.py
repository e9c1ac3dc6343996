"""Bucket the warp-state samples of an `ncu --page source --csv --print-source sass` dump by code segment (runs of
instructions with the same execution count): where does a kernel's time go - loop bodies, reductions, prologues?"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
keys = ['stall_wait', 'stall_long_sb', 'stall_short_sb', 'stall_math', 'stall_selected', 'stall_dispatch', 'stall_not_selected', 'stall_no_inst', 'stall_branch_resolving']
recs = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    recs.append((int(r[ix['Address']], 16), r[ix['Source']].strip(), int(r[ix['# Samples']] or 0), int(r[ix['Instructions Executed']] or 0), [int(r[ix[k]] or 0) for k in keys]))
recs.sort()
tot = sum(r[2] for r in recs)
seg = []; cur = None
for a, src, n, e, st in recs:
    if cur is None or cur['e'] != e:
        if cur: seg.append(cur)
        cur = {'e': e, 'n': 0, 'cnt': 0, 'wide': 0, 'st': [0] * len(keys), 'first': src}
    cur['n'] += n; cur['cnt'] += 1
    cur['st'] = [x + y for x, y in zip(cur['st'], st)]
    if 'IMAD.WIDE' in src: cur['wide'] += 1
seg.append(cur)
print('total samples', tot, '| columns:', ' '.join(k.replace('stall_', '') for k in keys))
for s in seg:
    if s['n'] / tot > float(sys.argv[2]) if len(sys.argv) > 2 else 0.01:
        print(f"exec={s['e']:>10} instrs={s['cnt']:4d} wide={s['wide']:3d} samples={100*s['n']/tot:5.1f}% | " + ' '.join(f"{100*x/tot:4.1f}" for x in s['st']) + f" | {s['first'][:40]}")
