"""ncu launch list (csv of gpu__time_duration.sum) -> per-kernel count / average / share.  usage: launch_shares.py in.csv [note]"""
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
hdr = rows[0]
ki, vi, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name'), hdr.index('Metric Unit')
agg = {}
for r in rows[1:]:
    if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
        continue
    us = float(r[vi].replace(',', '')) / (1e3 if r[ui] in ('ns', 'nsecond') else 1.0)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
if len(sys.argv) > 2:
    print('# ' + sys.argv[2])
print('# per-launch times under ncu are cold-cache and serialised: compare SHARES; total %.1f us over %d launches' % (tot, sum(v[0] for v in agg.values())))
for k, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-90s n=%4d  avg %9.1f us  share %5.1f%%' % (k[:90], n, us / n, 100 * us / tot))
