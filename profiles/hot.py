"""hot.py <ncu-rep> <rows per launch>: per-SASS-instruction executed counts / samples joined with source lines."""
import csv, subprocess, sys, collections, io, re
rep, rows_per = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
h = rows[hi]; ci = {n: i for i, n in enumerate(h)}
print([c for c in h][:6])
ins = []
for r in rows[hi+1:]:
    if len(r) <= ci['Instructions Executed']: continue
    try:
        ins.append((int(r[ci['Instructions Executed']]), int(r[ci['# Samples']] or 0), r[ci['Source']], int(r[ci['Thread Instructions Executed']])))
    except ValueError:
        pass
tot = sum(i[0] for i in ins); ts = sum(i[1] for i in ins)
print('total winst per row %.0f, samples %d' % (tot / rows_per, ts))
ops = collections.Counter()
for n, s, src, t in ins:
    tok = src.split()
    op = tok[1] if tok and tok[0].startswith('@') and len(tok) > 1 else (tok[0] if tok else '?')
    ops[op.split('.')[0]] += n
print([(o, round(c / rows_per)) for o, c in ops.most_common(30)])
print("--- top by samples")
for k in sorted(range(len(ins)), key=lambda k: -ins[k][1])[:40]:
    print('%5d  exec/row %7.1f  thr/inst %4.1f  %5.1f%%  %s' % (k, ins[k][0] / rows_per, ins[k][3] / max(1, ins[k][0]), 100 * ins[k][1] / ts, ins[k][2][:100]))
if len(sys.argv) > 3:
    pat = sys.argv[3]
    print("--- instructions matching", pat, "(with the 3 following)")
    for k in range(len(ins)):
        if re.search(pat, ins[k][2]):
            for j in range(k, min(k + 4, len(ins))):
                print('%5d  exec/row %7.1f  thr/inst %4.1f  %5.2f%%  %s' % (j, ins[j][0] / rows_per, ins[j][3] / max(1, ins[j][0]), 100 * ins[j][1] / ts, ins[j][2][:90]))
            print()
