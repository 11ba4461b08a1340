"""Per-instruction stall profile from `ncu --page source --csv`: python profiles/sass_hot.py file.csv [lo hi]
Prints cumulative sample share per contiguous code region and the hottest instructions with their
dominant stall reason."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h)]
# ncu lists the function twice when two launches are in the report: keep the first copy
addr0 = data[0][0]
dup = [i for i, r in enumerate(data) if r[0] == addr0]
if len(dup) > 1:
    data = data[:dup[1]]
col = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]


def f(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return 0.0


S = [f(r[col['# Samples']]) for r in data]
tot = sum(S)
print('instructions', len(data), 'samples', tot)
tot_by = {n: sum(f(r[col[n]]) for r in data) for n in stalls}
print('stall mix:', ' '.join('%s=%.1f%%' % (n[6:], 100 * v / tot) for n, v in sorted(tot_by.items(), key=lambda kv: -kv[1]) if v / tot > 0.01))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
W = 40
print('--- sample share per %d-instruction window' % W)
for a in range(lo, hi, W):
    sh = sum(S[a:a + W]) / tot
    if sh > 0.01:
        ex = f(data[a][col['Instructions Executed']])
        print('[%4d:%4d] %5.1f%%  exec=%6.0f  %s' % (a, a + W, 100 * sh, ex, data[a][1].strip()[:50]))
print('--- hottest instructions')
for i in sorted(range(lo, hi), key=lambda i: -S[i])[:45]:
    r = data[i]
    top = max(stalls, key=lambda n: f(r[col[n]]))
    print('%4d %5.2f%% exec=%6.0f %-14s %s' % (i, 100 * S[i] / tot, f(r[col['Instructions Executed']]), top[6:], r[1].strip()[:70]))
