"""Warp-stall totals and the hottest SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv`,
plus a few headline metrics from `--page raw --csv` (development aid; writes the text kept under profiles/).
   python tools/ncu_stall_summary.py source.csv raw.csv "title" > profiles/rNN_<kernel>_stalls.txt"""
import csv, sys

src, raw, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[h]
stall_cols = [i for i, c in enumerate(H) if c.startswith("stall_") and "(Not Issued)" not in c]
si = H.index("Source")
ni = H.index("# Samples")
tot = {H[i]: 0 for i in stall_cols}
inst = []
for r in rows[h + 1:]:
    if len(r) <= max(stall_cols):
        continue
    vals = {H[i]: int(r[i] or 0) for i in stall_cols}
    for k, v in vals.items():
        tot[k] += v
    inst.append((int(r[ni] or 0), r[0], " ".join(r[si].split()), vals))
print(f"# {title}")
all_s = sum(tot.values())
print("# warp-stall samples summed over the kernel (share of all samples)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"{k:24s} {v:9d}  {100.0 * v / all_s:5.1f} %")
print()
print("# hottest instructions: samples, address, SASS, dominant stall reasons")
for n, addr, sass, vals in sorted(inst, key=lambda t: -t[0])[:30]:
    top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(vals.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{n:7d}  {addr[-6:]}  {sass[:60]:60s}  {top}")
print()
R = list(csv.reader(open(raw)))
hr = next(i for i, r in enumerate(R) if r and r[0] == "ID")
names, units, vals = R[hr], R[hr + 1], R[hr + 2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]
print("# headline metrics of the same capture")
for w in want:
    if w in names:
        i = names.index(w)
        print(f"{w:60s} {vals[i]:>16s} {units[i]}")
