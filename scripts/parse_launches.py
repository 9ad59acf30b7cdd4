"""Prints an ncu launch list (--metrics gpu__time_duration.sum,smsp__inst_executed.sum --csv) per kernel launch."""
import csv
import sys

for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    byid = {}
    for r in rows[1:]:
        byid.setdefault(r[ii], {'k': r[ki][:48]})[r[mi]] = float(r[vi].replace(',', ''))
    print(f)
    for i, v in byid.items():
        t = v.get('gpu__time_duration.sum', 0) / 1e6
        n = v.get('smsp__inst_executed.sum', 0)
        print(f"{i:>3} {v['k']:48s} {t:8.3f} ms {n / 1e9:7.3f} G warp-inst {n * 32 / 2 ** 30:7.1f} inst/amp(30q)")
