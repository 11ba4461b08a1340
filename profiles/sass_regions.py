"""Region summary of an `ncu --page source --csv` (SASS) export:
python profiles/sass_regions.py file.csv [--top N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
data = [dict(zip(h, r)) for r in rows[2:] if len(r) == len(h)]


def f(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return 0.0


tot_i = sum(f(d['Instructions Executed']) for d in data)
tot_s = sum(f(d['# Samples']) for d in data)
print('SASS instrs', len(data), 'warp-instr executed', tot_i, 'samples', tot_s)
i = 0
while i < len(data):
    e = f(data[i]['Instructions Executed'])
    j, s, tot = i, 0, 0
    ops = {}
    while j < len(data) and abs(f(data[j]['Instructions Executed']) - e) <= 0.25 * max(e, 1):
        s += f(data[j]['# Samples'])
        tot += f(data[j]['Instructions Executed'])
        src = data[j]['Source'].split()
        op = src[1] if src[0].startswith('@') else src[0]
        ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + 1
        j += 1
    if tot / tot_i > 0.01 or s / tot_s > 0.01:
        top = sorted(ops.items(), key=lambda kv: -kv[1])[:8]
        print('SASS[%4d:%4d] n=%3d exec/instr=%8.0f inst%% %5.1f samp%% %5.1f  %s' %
              (i, j, j - i, e, 100 * tot / tot_i, 100 * s / tot_s, ' '.join('%s:%d' % kv for kv in top)))
    i = j
if '--top' in sys.argv:
    n = int(sys.argv[sys.argv.index('--top') + 1])
    for d in sorted(data, key=lambda d: -f(d['# Samples']))[:n]:
        print('%5.1f%%  %s' % (100 * f(d['# Samples']) / tot_s, d['Source'].strip()[:100]))
