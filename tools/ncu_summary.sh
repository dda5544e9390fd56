#!/bin/bash
# usage: tools/ncu_summary.sh gpurun_out/X.ncu-rep profiles/NAME   (run here, no GPU needed)
rep=$1; out=$2
ncu -i $rep --page details 2>/dev/null | grep -E "^\s+(void )?[a-z_]+<|^  [a-z_]+\(|Duration|Elapsed Cycles|SM Frequency|Executed Ipc Active|Issue Slots Busy|Registers Per Thread|Theoretical Occupancy|Achieved Occupancy|Avg. Active Threads Per Warp|Avg. Not Predicated|L1/TEX Hit Rate|L2 Hit Rate|DRAM Throughput|Memory Throughput|L2 Cache Throughput|L1/TEX Cache Throughput|Executed Instructions  |No Eligible|Warp Cycles Per Issued|Compute \(SM\) Throughput|Grid Size|Block Size" > ${out}_ncu_details.txt
ncu -i $rep --page raw --csv 2>/dev/null > /tmp/_raw.csv
python - "$out" <<'PY'
import csv, sys, json
out = sys.argv[1]
r = list(csv.reader(open('/tmp/_raw.csv')))
h = r[0]
rows = r[2:]
ki = h.index("Kernel Name")
def col(name): return h.index(name)
dr, dw, du = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
unit = r[1]
def tobytes(v, u):
    v = float(v); return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
tot = 0.0; lines = []
for x in rows:
    b = tobytes(x[dr], unit[dr]) + tobytes(x[dw], unit[dw]); tot += b
    lines.append("%s  dram read+write %.1f MB  duration %s %s" % (x[ki].split("(")[0], b / 1e6, x[du], unit[du]))
open(out + "_ncu_details.txt", "a").write("\n# per captured launch (ncu --set full, cold caches, serialised)\n" + "\n".join(lines) + "\n")
json.dump({"launches": len(rows), "dram_bytes_per_launch": tot / max(1, len(rows))}, open(out + "_traffic.json", "w"))
print(out, len(rows), "launches, dram bytes/launch %.1f MB" % (tot / max(1, len(rows)) / 1e6))
PY
ncu -i $rep --page source --print-source cuda,sass --csv > /tmp/_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/_src.csv 25 > ${out}_hot_lines.txt
