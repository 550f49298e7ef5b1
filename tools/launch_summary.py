"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(list)
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r: hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    v = float(d['Metric Value'].replace(',', ''))
    unit = d['Metric Unit']
    if unit == 'ns': v /= 1000
    elif unit == 'ms': v *= 1000
    agg[d['Kernel Name'][:70]].append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print('%-72s n=%4d total %10.1f us (%4.1f%%) mean %8.1f min %8.1f max %8.1f' % (k, len(v), sum(v), 100*sum(v)/tot, sum(v)/len(v), min(v), max(v)))
