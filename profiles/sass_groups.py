import csv, re, sys, collections
ncu_csv, disasm, kernel_pat, rows_per_launch = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
groups = eval(open(sys.argv[5]).read())   # list of (name, file, lo, hi)
rows = list(csv.reader(open(ncu_csv)))
h = rows[1]; ci = {n: i for i, n in enumerate(h)}
ins = [(int(r[ci['Instructions Executed']]), int(r[ci['Thread Instructions Executed']]), int(r[ci['# Samples']] or 0), r[ci['Source']]) for r in rows[2:] if len(r) > 5]
lines = open(disasm).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and re.search(kernel_pat, l))
cur = None; seq = []
for l in lines[start+1:]:
    if (l.startswith('.text.') or l.strip().startswith('.section')) and seq: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append(cur)
agg = collections.defaultdict(lambda: [0,0,0]); ops = collections.defaultdict(collections.Counter)
for k in range(min(len(ins), len(seq))):
    loc = seq[k] or ('?', 0)
    name = 'other'
    for (g, f, lo, hi) in groups:
        if loc[0] == f and lo <= loc[1] <= hi: name = g; break
    agg[name][0] += ins[k][0]; agg[name][1] += ins[k][1]; agg[name][2] += ins[k][2]
    op = ins[k][3].split()[0] if ins[k][3].split() else '?'
    if op.startswith('@'): op = ins[k][3].split()[1]
    ops[name][op.split('.')[0]] += ins[k][0]
tot = sum(v[0] for v in agg.values()); tots = sum(v[2] for v in agg.values())
print(f"total warp-instr {tot}  per row {tot/rows_per_launch:.0f}")
for name, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    top = ', '.join(f"{o}:{c/rows_per_launch:.0f}" for o, c in ops[name].most_common(8))
    print(f"{name:14s} {100*v[0]/tot:5.1f}% inst  {v[0]/rows_per_launch:7.0f} winst/row  thr/inst {v[1]/max(1,v[0]):4.1f}  {100*v[2]/max(1,tots):5.1f}% samples | {top}")
