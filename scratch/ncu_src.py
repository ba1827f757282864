"""aggregate the ncu source page: top SASS instructions by stall samples, and per-opcode totals."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
# find header row
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
def f(r, k):
    try: return float(r[idx[k]])
    except: return 0.0
tot_samples = sum(f(r, '# Samples') for r in data)
tot_inst = sum(f(r, 'Instructions Executed') for r in data)
print('total samples', tot_samples, 'warp-instr', tot_inst)
byop = collections.defaultdict(lambda: [0, 0])
for r in data:
    op = r[idx['Source']].split()[0] if r[idx['Source']].split() else '?'
    if op.startswith('@'): op = r[idx['Source']].split()[1]
    byop[op][0] += f(r, 'Instructions Executed'); byop[op][1] += f(r, '# Samples')
print('op, %instr, %samples')
for op, (n, s) in sorted(byop.items(), key=lambda kv: -kv[1][1])[:18]:
    print('%-12s %6.2f %6.2f' % (op, 100 * n / tot_inst, 100 * s / tot_samples))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('\ntop instructions by samples:')
top = sorted(data, key=lambda r: -f(r, '# Samples'))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    st = sorted(((f(r, s), s) for s in stalls), reverse=True)[:3]
    print('%s %5.2f%% exec=%d  %-50s %s' % (r[idx['Address']][-5:], 100 * f(r, '# Samples') / tot_samples, f(r, 'Instructions Executed'), r[idx['Source']][:50], ' '.join('%s=%d' % (s[6:], v) for v, s in st if v > 0)))
