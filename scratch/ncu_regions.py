"""samples / instructions by SASS address bucket (code-region view of an ncu source page)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
def f(r, k):
    try: return float(r[idx[k]])
    except: return 0.0
base = min(int(r[idx['Address']], 16) for r in data)
tot = sum(f(r, '# Samples') for r in data)
buckets = collections.OrderedDict()
for r in data:
    a = (int(r[idx['Address']], 16) - base) // B
    b = buckets.setdefault(a, [0, 0, 0, collections.Counter()])
    b[0] += f(r, '# Samples'); b[1] += f(r, 'Instructions Executed'); b[2] += 1
    for st in ('stall_no_inst', 'stall_short_sb', 'stall_long_sb', 'stall_wait', 'stall_not_selected', 'stall_dispatch', 'stall_branch_resolving'):
        b[3][st[6:]] += f(r, st)
print('code size: %d instrs = %.1f KB' % (len(data), len(data) * 16 / 1024))
for a, (s, n, c, st) in buckets.items():
    if s / tot > 0.01:
        print('off %6d KB: %5.1f%% samples, %6.1fM warp-instr, %4d static instrs | %s' % (a * B // 1024, 100 * s / tot, n / 1e6, c, ' '.join('%s=%.0f%%' % (k, 100 * v / max(s, 1)) for k, v in st.most_common(4))))
