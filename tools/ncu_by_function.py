"""Aggregate 'Instructions Executed' of an `ncu --page source --csv --print-source cuda,sass` export per source
function and per line:   ncu_by_function.py EXPORT.csv SOURCE.cu [TOPLINES]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2]
iA, iS, iI = 2, hdr.index("# Samples"), hdr.index("Instructions Executed")
agg = {}
for r in rows[3:]:
    if len(r) > iI and r[iA] == '-' and r[0].isdigit():
        agg[int(r[0])] = (int(r[iI]), int(r[iS]), r[1])
tot = sum(v[0] for v in agg.values())
print('total inst', tot, 'samples', sum(v[1] for v in agg.values()))
src = open(sys.argv[2]).read().splitlines()
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r'^(?:template.*)?(?:static\s+)?__(?:device|global)__.*?\b(\w+)\s*\(', l)
    if m:
        funcs.append((i, m.group(1)))
funcs.append((len(src) + 1, 'END'))
out = []
for (a, n), (b, _) in zip(funcs, funcs[1:]):
    s = sum(v[0] for k, v in agg.items() if a <= k < b)
    sm = sum(v[1] for k, v in agg.items() if a <= k < b)
    if s:
        out.append((s, sm, n, a))
for s, sm, n, a in sorted(out, reverse=True):
    print('%9d %5.1f%%  samples %5d  %s (line %d)' % (s, 100 * s / tot, sm, n, a))
print()
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(k, v[0], v[1], v[2][:110])
