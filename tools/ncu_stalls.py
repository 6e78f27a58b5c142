"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('total samples', tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: 0 for s in stalls}
for r in data:
    for s in stalls:
        v = r[ix[s]]
        if v:
            agg[s] += int(v)
print({k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
top = sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:n]
for r in top:
    st = {s[6:]: int(r[ix[s]]) for s in stalls if r[ix[s]] and int(r[ix[s]]) > 0}
    print(r[ix['Address']][-5:], r[ix['# Samples']], r[ix['Source']][:80], st)
