"""Print kernel name / duration (ms) pairs from an `ncu --metrics gpu__time_duration.sum --csv --log-file` launch list."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    if r[h.index('Metric Name')] == 'gpu__time_duration.sum':
        name = r[h.index('Kernel Name')]
        name = name.replace('void ', '').replace('bls::', '').split('(')[0]
        print(f"{name[:60]:60s} {float(r[h.index('Metric Value')]) / 1e6:10.3f} ms")
