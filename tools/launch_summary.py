"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum [+ dram bytes]) per kernel and grid."""
import csv, collections, re, sys
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki, vi, ui, gi, mi, idi = (H.index(n) for n in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Grid Size', 'Metric Name', 'ID'))
per = collections.defaultdict(dict)
for r in data:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    if u == 'ns': v /= 1e3
    elif u == 'ms': v *= 1e3
    elif u == 'Kbyte': v /= 1e3
    elif u == 'byte': v /= 1e6
    elif u == 'Gbyte': v *= 1e3
    per[r[idi]][r[mi]] = v
    per[r[idi]]['name'] = re.sub(r'\(.*', '', r[ki]).replace('void ', '') ; per[r[idi]]['grid'] = r[gi]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k, d in per.items():
    a = agg[(d['name'], d['grid'])]
    a[0] += 1; a[1] += d.get('gpu__time_duration.sum', 0); a[2] += d.get('dram__bytes_read.sum', 0); a[3] += d.get('dram__bytes_write.sum', 0)
tot = sum(a[1] for a in agg.values())
print(f"{'total us':>10} {'share':>6} {'n':>5} {'avg us':>8} {'rd MB':>8} {'wr MB':>8} {'GB/s':>7}  kernel grid")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (a[2] + a[3]) / a[1] * 1e3 if a[1] else 0
    print(f"{a[1]:10.1f} {100*a[1]/tot:5.1f}% {a[0]:5d} {a[1]/a[0]:8.1f} {a[2]/a[0]:8.1f} {a[3]/a[0]:8.1f} {gbs:7.0f}  {k[0]} {k[1]}")
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
