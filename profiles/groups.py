"""groups.py <ncu-rep> <rows per launch> <cubin kernel regex>: executed warp instructions and stall samples of k_fused_flow per function group (line ranges of the round-2 sources)."""
import csv, subprocess, sys, collections, io, re, os, glob
rep, rows_per, pat = sys.argv[1], float(sys.argv[2]), sys.argv[3]
lib = "/root/repo/structure-light-reconstructor_b200/libslr_b200.so"
os.makedirs("/tmp/cub2", exist_ok=True)
for f in glob.glob("/tmp/cub2/*.cubin"): os.remove(f)
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd="/tmp/cub2", capture_output=True)
cub = [f for f in glob.glob("/tmp/cub2/*.cubin") if "k_fused_flow" in f][0]
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
def group(loc):
    if not loc: return 'other'
    f, ln = loc
    if f == 'k_fused_common.cuh':
        if 223 <= ln <= 263 or 345 <= ln <= 360: return 'chain walk'
        if 264 <= ln <= 330 or 87 <= ln <= 104: return 'insert (probe, claim, minimum, bucket filing)'
        if 108 <= ln <= 217: return 'decode: plane loads, shadow test'
        if ln <= 81: return 'job control (bulk copies)'
    if f == 'slr_device.cuh':
        if 188 <= ln <= 202: return 'decode: byte differences (IDP.4A)'
        if 203 <= ln <= 256: return 'decode: table lookups'
        if 257 <= ln <= 286: return 'decode: heterodyne'
        if 355 <= ln <= 445: return 'reprojection (fp64) + exact narrowing'
        if ln <= 120: return 'job control (mbarrier waits)'
    if f.startswith('device_atomic'): return 'insert (probe, claim, minimum, bucket filing)'
    if f == 'k_fused_flow.cu':
        if 378 <= ln <= 416: return 'stores'
        if 345 <= ln <= 377: return 'query set-up, map loads'
        if 300 <= ln <= 312: return 'decode: shuffle, park'
        return 'job control'
    if f.startswith('sm_'): return 'job control (shuffles, waits)'
    return 'other ' + f
agg = collections.defaultdict(lambda: [0, 0])
for k in range(min(len(ins), len(seq))):
    g = group(seq[k]); agg[g][0] += ins[k][0]; agg[g][1] += ins[k][1]
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print('total %.0f warp instructions per row' % (tot / rows_per))
for g, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%-50s %7.0f winst/row %5.1f%%  samples %5.1f%%' % (g, v[0]/rows_per, 100*v[0]/tot, 100*v[1]/ts))
