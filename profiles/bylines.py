"""bylines.py <ncu-rep> <rows per launch> <cubin kernel regex>: executed warp-instructions and stall samples per source line."""
import csv, subprocess, sys, collections, io, re, os, glob
rep, rows_per, pat = sys.argv[1], float(sys.argv[2]), sys.argv[3]
lib = os.environ.get("LIB", "/root/repo/structure-light-reconstructor_b200/libslr_b200.so")
os.makedirs("/tmp/cub", exist_ok=True)
for f in glob.glob("/tmp/cub/*.cubin"): os.remove(f)
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd="/tmp/cub", capture_output=True)
cub = [f for f in glob.glob("/tmp/cub/*.cubin") if os.environ.get("CUBIN", "k_fused") in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r)
h = rows[hi]; ci = {n: i for i, n in enumerate(h)}
ins = []
for r in rows[hi+1:]:
    try: ins.append((int(r[ci['Instructions Executed']]), int(r[ci['# Samples']] or 0), r[ci['Source']]))
    except (ValueError, IndexError): pass
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and re.search(pat, l))
cur = None; seq = []
for l in dis[start+1:]:
    if (l.startswith('.text.') or l.strip().startswith('.section')) and seq: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l): seq.append(cur)
print(len(ins), len(seq))
agg = collections.defaultdict(lambda: [0, 0])
for k in range(min(len(ins), len(seq))):
    loc = seq[k] or ('?', 0)
    agg[loc][0] += ins[k][0]; agg[loc][1] += ins[k][1]
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
src = {}
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    if f not in src:
        pth = os.path.join("/root/repo/structure-light-reconstructor_b200/csrc", f)
        src[f] = open(pth).read().splitlines() if os.path.exists(pth) else []
    text = src[f][ln-1].strip()[:90] if 0 < ln <= len(src[f]) else ''
    print('%-16s %4d  %7.0f winst/row %5.1f%%  samples %5.1f%% | %s' % (f, ln, v[0]/rows_per, 100*v[0]/tot, 100*v[1]/ts, text))
