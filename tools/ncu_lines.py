"""Map the per-instruction samples of an `ncu --page source --csv` export (SASS view) to CUDA source lines through
`nvdisasm -g` line info of the same cubin, and print the top source lines by stall samples.
usage: python tools/ncu_lines.py <source.csv> <nvdisasm_-g_output.txt> [top]"""
import csv, re, sys, collections

def main():
    src_csv, dis, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
    # offsets -> (line, inlined-at chain ignored: innermost location)
    off2line, cur = {}, None
    for ln in open(dis, errors="replace"):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = int(m.group(2)); continue
        m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
        if m and cur is not None:
            off2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(src_csv)))
    h = rows[1]
    ia, isamp, iinst = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    base = int(rows[2][ia], 16)
    per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot = 0
    for r in rows[2:]:
        off = int(r[ia], 16) - base
        line = off2line.get(off, -1)
        s = int(r[isamp] or 0); tot += s
        per[line][0] += s; per[line][1] += int(r[iinst] or 0)
        for i in stall_cols:
            v = int(r[i] or 0)
            if v: per[line][2][h[i]] += v
    print("total samples", tot)
    for line, (s, n, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"line {line:5d}  samples {s:7d} ({100*s/tot:5.1f}%)  inst {n:10d}  " + " ".join(f"{k[6:]}={v}" for k, v in st.most_common(3)))

main()
