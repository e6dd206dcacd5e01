"""Top stalled SASS instructions of an ncu source page:  ncu -i X.ncu-rep --page source --csv | python tools/ncu_top_stalls.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
ci = {k: i for i, k in enumerate(h)}
data, tot = [], 0
for r in rows[hi + 1:]:
    try:
        s = int(r[ci['# Samples']])
    except Exception:
        continue
    tot += s
    data.append((s, r))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
stallcols = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
agg = {k: sum(int(r[ci[k]] or 0) for _, r in data) for k in stallcols}
print('total samples', tot, 'instructions', len(data))
print('by reason', sorted(((v, k[6:]) for k, v in agg.items() if v), reverse=True))
for s, r in sorted(data, key=lambda x: -x[0])[:n]:
    st = sorted([(int(r[ci[k]] or 0), k[6:]) for k in stallcols], reverse=True)[:2]
    print(s, r[ci['Source']][:100], st)
