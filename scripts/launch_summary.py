"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the LAST step."""
import collections, csv, re, sys
path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
detail = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr, data = rows[h], rows[h + 1:]
ki, vi, gi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size'), hdr.index('Metric Unit')
step = data[len(data) - len(data) // nsteps:]
agg = collections.OrderedDict()
for r in step:
    name = re.sub(r'\(.*', '', r[ki])
    if 'ew_kernel' in name:
        name = 'ttb::ew_kernel<...>'
    t = float(r[vi].replace(',', '')) * (1e-3 if r[ui] in ('ns', 'nsecond') else 1)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{len(step)} launches in the last step, {tot:.1f} us total (serialised, cold-cache)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}% {c:4d}x  {k[:100]}")
if detail:
    print()
    for r in step:
        if detail in r[ki]:
            print(re.sub(r'\(ttb.*', '', r[ki])[:60], r[gi], r[vi], r[ui])
