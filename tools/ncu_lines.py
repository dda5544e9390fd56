"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv; python tools/ncu_lines.py x.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
fname = ""
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
        continue
    def g(k):
        try:
            return float(r[hdr.index(k)])
        except ValueError:
            return 0.0
    out.append((g("Instructions Executed"), g("Thread Instructions Executed"), g("# Samples"), fname, r[0], r[1].strip()[:100]))
ti = sum(o[0] for o in out); ts = sum(o[2] for o in out); tt = sum(o[1] for o in out)
print("total warp-inst %.3g  thread-inst %.3g  avg threads/inst %.2f  samples %d" % (ti, tt, tt / max(ti, 1), ts))
print("by file:")
files = {}
for o in out:
    f = files.setdefault(o[3], [0, 0, 0]); f[0] += o[0]; f[1] += o[1]; f[2] += o[2]
for k, f in sorted(files.items(), key=lambda kv: -kv[1][2]):
    print("  %-22s inst %5.1f%%  samples %5.1f%%  thr/inst %5.1f" % (k, 100 * f[0] / ti, 100 * f[2] / ts, f[1] / max(f[0], 1)))
print("top lines by stall samples:")
for o in sorted(out, key=lambda o: -o[2])[:top]:
    print("%5.1f%% inst %5.1f%% smp thr %4.1f | %s:%s %s" % (100 * o[0] / ti, 100 * o[2] / ts, o[1] / max(o[0], 1), o[3], o[4], o[5]))
