"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys


def main(path, skip_until=None):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'^void ', '', r['Kernel Name'])
        name = re.sub(r'\(.*', '', name)
        grid = r['Grid Size']
        key = (name[:90], grid)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r['Metric Value']) / 1e6
    tot = sum(v[1] for v in agg.values())
    print(f'# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} ms total (cold-cache, serialised)')
    print(f'# {"ms total":>10} {"launches":>8} {"ms/launch":>10} {"share":>6}  kernel  grid')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{v[1]:12.2f} {v[0]:8d} {v[1] / v[0]:10.3f} {100 * v[1] / tot:5.1f}%  {k[0]}  {k[1]}')


if __name__ == '__main__':
    main(sys.argv[1])
