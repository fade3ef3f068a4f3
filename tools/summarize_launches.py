"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ik, iv, iu, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
agg = collections.OrderedDict()
tot = 0.
for row in r:
    name = row[ik]
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'\(.*$', '', name)
    name = name.replace('fused::', '').replace('tail::', '')
    if 'elementwise' in name and 'lambda' in row[ik]:
        m = re.search(r'f2d_(\w+)', row[ik])
        name = 'k_elementwise:' + (m.group(1) if m else '?')
    if 'k_reduce' in name:
        m = re.search(r'f2d_(\w+)', row[ik])
        name = re.sub(r'<.*', '', name) + ':' + (m.group(1) if m else '?')
    v = float(row[iv].replace(',', ''))
    u = row[iu]
    v *= {'ns': 1e-3, 'us': 1., 'usecond': 1., 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3}.get(u, 1.)
    big = row[gi]
    key = (name, big)
    a = agg.setdefault(key, [0, 0.])
    a[0] += 1
    a[1] += v
    tot += v
print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
byname = collections.OrderedDict()
for (name, g), a in agg.items():
    b = byname.setdefault(name, [0, 0.])
    b[0] += a[0]
    b[1] += a[1]
print("---- by kernel")
for name, a in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print("%-60s n=%5d total=%9.1f us %5.1f%%" % (name[:60], a[0], a[1], 100*a[1]/tot))
print("---- top (kernel, grid)")
for (name, g), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%-50s grid=%-16s n=%4d avg=%8.2f us total=%9.1f us %5.1f%%" % (name[:50], g, a[0], a[1]/a[0], a[1], 100*a[1]/tot))
