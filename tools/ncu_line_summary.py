"""Per-source-line warp-stall samples of one kernel (development aid).
Joins `ncu -i X.ncu-rep --page source --csv` (per SASS instruction, absolute addresses) with the
line table of the same kernel from `nvdisasm -g -fun <symbol index> <cubin>` (offsets + //## File lines).
   cuobjdump -xelf all lib/libimomd_b200.so; cuobjdump -elf X.cubin | grep '.text.*<kernel>'   # info column = index
   nvdisasm -g -fun 0x<index> X.cubin > k.sass
   python tools/ncu_line_summary.py source.csv k.sass "title" [top] [source root] > profiles/rNN_<kernel>_lines.txt"""
import csv, re, sys, collections

src, sass, title = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
src_root = sys.argv[5] if len(sys.argv) > 5 else "/root/repo"   # the sources of the build that was profiled
line_of = {}
cur = None
started = False
for ln in open(sass, errors="replace"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        started = True
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and started:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src, errors="replace")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[h]
stall_cols = [i for i, c in enumerate(H) if c.startswith("stall_") and "(Not Issued)" not in c]
ni = H.index("# Samples")
ei = H.index("Instructions Executed")
base = int(rows[h + 1][0], 16)
per = collections.defaultdict(lambda: [0, collections.Counter(), 0])
tot = collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= max(stall_cols):
        continue
    off = int(r[0], 16) - base
    key = line_of.get(off, ("?", 0))
    n = int(r[ni] or 0)
    per[key][0] += n
    per[key][2] += int(r[ei] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v:
            per[key][1][H[i][6:]] += v
            tot[H[i][6:]] += v
alls = sum(tot.values())
print(f"# {title}")
print(f"samples: {alls};  " + ", ".join(f"{k} {100.0 * v / alls:.1f}%" for k, v in tot.most_common(9)))
text = {}
def source_line(f, l):
    if f not in text:
        try:
            import glob
            p = glob.glob(f"{src_root}/**/{f}", recursive=True)
            text[f] = open(p[0], errors="replace").read().splitlines() if p else []
        except OSError:
            text[f] = []
    t = text[f]
    return t[l - 1].strip()[:90] if 0 < l <= len(t) else ""
for key, (n, c, ex) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100.0 * n / alls:5.2f}%  {key[0]}:{key[1]:<5d} {source_line(*key):90s} [{', '.join(f'{k} {v}' for k, v in c.most_common(2))}] inst {ex}")
