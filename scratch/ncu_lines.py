"""attribute ncu source-page samples to CUDA source lines via nvdisasm -g line info.
usage: ncu_lines.py <src.csv> <cubin> <kernel-mangled-substring> [topN]"""
import csv, sys, re, subprocess, collections
src, cubin, kern = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.split('\n')
start = [i for i, l in enumerate(dis) if l.startswith('.text.') and kern in l and l.rstrip().endswith(':')][0]
line_of = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith('.text.') or l.startswith('//--------------------- .text'):
        if line_of: break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(\S.*?);', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
base = min(int(r[idx['Address']], 16) for r in data)
def f(r, k):
    try: return float(r[idx[k]])
    except: return 0.0
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
tot = 0
for r in data:
    off = int(r[idx['Address']], 16) - base
    key = line_of.get(off, ('?', 0))
    a = agg[key]; s = f(r, '# Samples'); a[0] += s; a[1] += f(r, 'Instructions Executed'); tot += s
    for st in ('stall_no_inst', 'stall_short_sb', 'stall_long_sb', 'stall_wait', 'stall_not_selected', 'stall_dispatch', 'stall_branch_resolving', 'stall_math'):
        a[2][st[6:]] += f(r, st)
print('total samples', tot)
for key, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print('%-22s:%-4d %5.1f%% samples %7.1fM instr | %s' % (key[0], key[1], 100 * s / tot, n / 1e6, ' '.join('%s=%.0f%%' % (k, 100 * v / max(s, 1)) for k, v in st.most_common(3))))
