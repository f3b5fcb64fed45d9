"""hot SASS lines of one kernel from an .ncu-rep (source page): python tools/ncu_hot.py rep kernel_regex [min_pct]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
for i, r in enumerate(rows):
    if 'Source' in r and '# Samples' in r:
        hdr, start = r, i
        break
si, k, ie = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
seen, data = set(), []
for r in rows[start + 1:]:
    if len(r) <= k or r[0] in seen:
        continue
    seen.add(r[0])
    try:
        data.append((float(r[k]), r[si], r[ie]))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print(len(data), 'instructions,', tot, 'samples')
acc = 0
for i, (v, s, n) in enumerate(data):
    acc += v
    if v / tot * 100 >= minp or any(x in s for x in ('BAR.SYNC', 'UBLKCP', 'TRYWAIT')):
        print(f'{i:5d} {v / tot * 100:6.2f}% cum {acc / tot * 100:5.1f}% n={n:>9s} {s[:100]}')
