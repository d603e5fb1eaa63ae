#!/bin/bash
# Development aid: where does the pseudo-tree merge (k_attach) spend its time on the bug-trap grid?
#   1. plain run with the facade's stage timers,  2. ncu launch list of the k_attach launches,
#   3. one full capture (source counters) of the LONGEST k_attach launch.
# Everything lands in gpurun_out/ with the prefix given as $1 (default r02_attach).
P=${1:-r02_attach}
W=${2:-1131}
IT=${3:-1000}
export IMOMD_NO_HOST_CERTIFY=1
IMOMD_FACADE_TIMING=1 python tools/bugtrap_probe.py $W $IT > gpurun_out/${P}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_attach --csv \
    --log-file gpurun_out/${P}_launches.csv python tools/bugtrap_probe.py $W $IT > gpurun_out/${P}_ncu1.log 2>&1
idx=$(python - "$P" <<'E'
import csv, sys
rows = list(csv.reader(open(f"gpurun_out/{sys.argv[1]}_launches.csv", errors="replace")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[h]
vi, ni = H.index("Metric Value"), H.index("Metric Name")
best, bi, n = -1.0, 0, 0
for r in rows[h + 1:]:
    if len(r) > vi and r[ni] == "gpu__time_duration.sum":
        v = float(r[vi].replace(",", ""))
        if v > best:
            best, bi = v, n
        n += 1
print(bi)
E
)
echo "longest k_attach launch: #$idx" >> gpurun_out/${P}_plain.log
ncu --set full --import-source on --clock-control none -k regex:k_attach -s $idx -c 1 -f -o gpurun_out/${P} \
    python tools/bugtrap_probe.py $W $IT > gpurun_out/${P}_ncu2.log 2>&1
ncu -i gpurun_out/${P}.ncu-rep --page raw --csv > gpurun_out/${P}_raw.csv
ncu -i gpurun_out/${P}.ncu-rep --page source --csv > gpurun_out/${P}_source.csv
ncu -i gpurun_out/${P}.ncu-rep --page source --print-source cuda --csv > gpurun_out/${P}_cuda.csv
ls -la gpurun_out/${P}*
