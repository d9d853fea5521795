"""Summarise an ncu source-page CSV (`ncu -i X.ncu-rep --page source --csv`): stall mix and the hottest instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
pairs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('total samples', tot, 'sass rows', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print('  %-24s %7d %5.1f%%' % (h, v, 100 * v / tot))
print('warp-instructions per unit', sum(int(r[ix['Instructions Executed']] or 0) for r in data) / pairs)
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    st = {h: int(r[ix[h]] or 0) for h in stalls}
    print(r[ix['# Samples']], r[ix['Instructions Executed']], max(st, key=st.get), r[ix['Source']][:100])
