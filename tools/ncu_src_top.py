"""Top stalled SASS instructions of one kernel from an `ncu --page source --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def num(v):
    try: return int(float(v))
    except ValueError: return 0
data = [r for r in rows[2:] if len(r) == len(hdr) and r[ix['# Samples']] != '# Samples']
tot = sum(num(r[ix['# Samples']]) for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(num(r[ix[h]]) for r in data) for h in stalls}
print('total samples', tot, sorted(agg.items(), key=lambda x: -x[1])[:9])
for r in sorted(data, key=lambda r: -num(r[ix['# Samples']]))[:n]:
    st = sorted(((h, num(r[ix[h]])) for h in stalls), key=lambda x: -x[1])[:2]
    print('%6d %5.1f%%  %-84s %s ex=%s' % (num(r[ix['# Samples']]), 100.0 * num(r[ix['# Samples']]) / max(tot, 1), r[ix['Source']][:84], st, r[ix['Instructions Executed']]))
