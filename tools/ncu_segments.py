"""Aggregate the stall samples of an `ncu --page source --csv --print-source sass` export by
barrier-delimited segment of the kernel (usage: ncu_segments.py export.csv)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ci['# Samples']]) for r in data)
seg = []


def fresh(i):
    return {'start': i, 'n': 0, 's': 0, 'st': {h: 0 for h in stalls}, 'ops': {}}


cur = fresh(0)
for i, r in enumerate(data):
    src = r[ci['Source']].strip()
    op = (src.split()[1] if src.startswith('@') else src.split()[0]).split('.')[0]
    cur['n'] += 1
    cur['s'] += int(r[ci['# Samples']])
    for h in stalls:
        cur['st'][h] += int(r[ci[h]] or 0)
    if op in ('LDG', 'STG', 'LDGSTS', 'LDS', 'STS', 'BAR', 'LDL', 'STL', 'CALL', 'MUFU'):
        cur['ops'][op] = cur['ops'].get(op, 0) + 1
    if op == 'BAR':
        cur['end'] = i + 1
        seg.append(cur)
        cur = fresh(i + 1)
cur['end'] = len(data)
seg.append(cur)
print(rows[0][1], '| samples', tot, '| SASS instructions', len(data))
for s in seg:
    top = sorted(s['st'].items(), key=lambda kv: -kv[1])[:4]
    print(f"{s['start']:5d}-{s['end']:5d} n={s['n']:4d} samp={100 * s['s'] / tot:5.1f}%  {dict(s['ops'])}  " +
          ' '.join(f"{k[6:]}={100 * v / max(1, s['s']):.0f}%" for k, v in top))
print({h[6:]: round(100 * sum(int(r[ci[h]] or 0) for r in data) / tot, 1) for h in stalls})
