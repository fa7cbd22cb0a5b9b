"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump: stall samples per CUDA
source line (development aid).  usage: ncu_hot_lines.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, cur_line, cur_src = None, None, None
agg = {}
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        si = hdr.index('# Samples')
        ii = hdr.index('Instructions Executed')
        continue
    if hdr is None or len(r) <= si:
        continue
    if r[0] not in ('', '-'):
        cur_line, cur_src = r[0], r[1]
        continue
    if r[2] in ('', '-', '...'):
        continue
    try:
        s = float(r[si] or 0)
        n = float(r[ii] or 0)
    except ValueError:
        continue
    k = (cur_file, cur_line, cur_src)
    a = agg.setdefault(k, [0.0, 0.0])
    a[0] += s
    a[1] += n
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print("total samples", tot, "total warp instructions", toti)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%s  %s" % (100 * a[0] / tot, 100 * a[1] / toti, k[0], k[1], (k[2] or '')[:100]))
