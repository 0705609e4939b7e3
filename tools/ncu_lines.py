#!/usr/bin/env python3
"""Join an ncu SASS source page (csv) with nvdisasm --print-line-info output to get per-CUDA-line instruction
counts and stall samples.  usage: ncu_lines.py <sass.csv> <nvdisasm.txt> <kernel mangled substring> [topN]"""
import csv, re, sys, collections
sass_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
# nvdisasm: sequence of instructions with preceding line markers
lines = open(dis).read().split('\n')
start = None
for k, l in enumerate(lines):
    if l.startswith('.text.') and kern in l:
        start = k; break
assert start is not None
seq = []  # (file,line) per instruction in order
cur = ('?', 0)
inl = ''
for l in lines[start + 1:]:
    if l.startswith('.text.') or l.startswith('//-----'):
        if seq: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        seq.append(cur)
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
iI = hdr.index('Instructions Executed'); iS = hdr.index('# Samples'); iT = hdr.index('Thread Instructions Executed')
cols = {n: hdr.index(n) for n in ['stall_no_inst', 'stall_long_sb', 'stall_barrier', 'stall_wait', 'stall_short_sb', 'stall_branch_resolving'] if n in hdr}
data = rows[2:]
print('sass rows', len(data), 'disasm instr', len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
n = min(len(data), len(seq))
for k in range(n):
    r = data[k]
    key = seq[k]
    a = agg[key]
    a[0] += int(r[iI] or 0); a[1] += int(r[iS] or 0); a[2] += int(r[iT] or 0)
    for nme, ix in cols.items():
        a[3][nme] += int(r[ix] or 0)
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print('total warp inst %d samples %d' % (tot, tots))
items = sorted(agg.items(), key=lambda kv: -kv[1][1])
src_cache = {}
def src(f, ln):
    import os
    p = '/root/repo/rnaelem_b200/csrc/' + f
    if p not in src_cache:
        src_cache[p] = open(p).read().split('\n') if os.path.exists(p) else []
    L = src_cache[p]
    return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ''
for (f, ln), a in items[:top]:
    st = ' '.join('%s=%d' % (k.replace('stall_', ''), v) for k, v in a[3].most_common(3))
    print('%5.2f%% smp %5.2f%% inst lanes %4.1f | %s:%d | %s | %s' % (100 * a[1] / tots, 100 * a[0] / tot, a[2] / max(a[0], 1), f, ln, st, src(f, ln)))

# ---- per-function aggregation (line ranges of dp_lin.cuh given as name:lo-hi on argv[5:])
if len(sys.argv) > 5:
    rng = []
    for sp in sys.argv[5:]:
        nm, r = sp.split(':'); lo, hi = r.split('-'); rng.append((nm, int(lo), int(hi)))
    fa = collections.defaultdict(lambda: [0, 0, 0])
    for (f, ln), a in agg.items():
        key = f
        if f == 'dp_lin.cuh':
            for nm, lo, hi in rng:
                if lo <= ln <= hi: key = nm; break
        fa[key][0] += a[0]; fa[key][1] += a[1]; fa[key][2] += a[2]
    print('--- by region')
    for k, v in sorted(fa.items(), key=lambda kv: -kv[1][1]):
        print('%-28s %5.2f%% smp %5.2f%% inst  lanes %4.1f' % (k, 100 * v[1] / tots, 100 * v[0] / tot, v[2] / max(v[0], 1)))
