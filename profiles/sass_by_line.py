"""Join ncu per-SASS-instruction counts (source page, sass view) with nvdisasm -g line info; aggregate by CUDA source line."""
import csv, re, sys, collections
ncu_csv, disasm, kernel_pat = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(ncu_csv)))
h = rows[1]
ci = {n: i for i, n in enumerate(h)}
ins = [(r[ci['Source']].strip(), int(r[ci['Instructions Executed']]), int(r[ci['Thread Instructions Executed']]), int(r[ci['# Samples']] or 0)) for r in rows[2:] if len(r) > 5]
# parse disasm for the kernel
lines = open(disasm).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('.text.') and re.search(kernel_pat, l))
cur = None; seq = []
for l in lines[start+1:]:
    if l.startswith('.text.') or l.strip().startswith('.section'): 
        if seq: break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)), m.group(3)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((m.group(2).strip(), cur))
print('ncu instrs', len(ins), 'disasm instrs', len(seq), file=sys.stderr)
agg = collections.Counter(); agg_s = collections.Counter()
n = min(len(ins), len(seq))
tot = sum(i[1] for i in ins)
for k in range(n):
    loc = seq[k][1]
    key = (loc[0], loc[1]) if loc else ('?', 0)
    agg[key] += ins[k][1]; agg_s[key] += ins[k][3]
src = {}
for (f, ln), c in agg.most_common(45):
    if f not in src:
        try: src[f] = open('/root/repo/structure-light-reconstructor_b200/csrc/' + f).read().splitlines()
        except Exception: src[f] = []
    text = src[f][ln-1].strip()[:90] if 0 < ln <= len(src[f]) else ''
    print(f"{100*c/tot:5.1f}% inst {100*agg_s[(f,ln)]/max(1,sum(agg_s.values())):5.1f}% samp  {f}:{ln:4d}  {text}")
