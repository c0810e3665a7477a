"""Aggregate the warp-state samples / executed instructions of an `ncu --set full --import-source on` capture by source
function and by source line of mss_kernels.cuh.
usage:  ncu -i X.ncu-rep --page source --csv > src.csv
        cuobjdump -xelf all ms_slam_b200/csrc/libmss.so && nvdisasm -g -c mss_engine.sm_100a.cubin > sass.txt
        python tools/ncu_functions.py src.csv sass.txt > profiles/<name>.txt"""
import csv, re, sys, collections, os

src_csv, dis = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = open(dis, errors='replace').read()
i = txt.index('.text._ZN3mss21mss_persistent_kernelENS_6ParamsE:')
j = txt.find('\n//--------------------- .text.', i)
off2line, cur = {}, None
for ln in txt[i:j if j > 0 else None].splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = int(m.group(2)); continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m and cur is not None:
        off2line[int(m.group(1), 16)] = cur
src = open(os.path.join(ROOT, 'ms_slam_b200/csrc/mss_kernels.cuh')).read().splitlines()
starts = []
for n, l in enumerate(src, 1):
    m = re.match(r'(?:template <[^>]*>\s*)?__(?:device|global)__ .*?(\w+)\(', l)
    if m:
        starts.append((n, m.group(1)))

def fn(line):
    name = '?'
    for n, nm in starts:
        if n <= line: name = nm
        else: break
    return name

rows = list(csv.reader(open(src_csv)))
h = rows[1]
ia, isamp, iinst = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [k for k, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
base = int(rows[2][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
ts = ti = 0
for r in rows[2:]:
    line = off2line.get(int(r[ia], 16) - base, -1)
    s, n = int(r[isamp] or 0), int(r[iinst] or 0)
    ts += s; ti += n
    for d, k in ((agg, fn(line)), (per, line)):
        d[k][0] += s; d[k][1] += n
        for c in stall_cols:
            v = int(r[c] or 0)
            if v: d[k][2][h[c]] += v
print("ncu --set full, mss_persistent_kernel, bench.py default launch (128 c2 windows): warp-state samples and executed warp")
print("instructions by function (innermost inlined location); total samples %d, total instructions %.0fM\n" % (ts, ti / 1e6))
for k, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{k:28s} samples {100*s/ts:5.1f}%  inst {100*n/ti:5.1f}%  top stalls: " + " ".join(f"{a[6:]}={100*b/max(s,1):.0f}%" for a, b in st.most_common(3)))
print("\nby source line (top 25)")
for line, (s, n, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:25]:
    code = src[line - 1].strip()[:100] if 0 < line <= len(src) else ''
    print(f"{line:5d} {100*s/ts:5.1f}%  " + " ".join(f"{a[6:]}={b}" for a, b in st.most_common(2)) + "  | " + code)
