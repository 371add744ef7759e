"""Summarise an `ncu --page source --csv` dump: samples and executed instructions per
segment between barriers / branch targets, plus the top SASS lines.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_hot.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_s = sum(int(r[ci["# Samples"]]) for r in data)
tot_i = sum(int(r[ci["Instructions Executed"]]) for r in data)
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
seg_s = seg_i = seg_f = 0
start = 0
print("segments (split at BAR / WARPSYNC / EXIT):")
for k, r in enumerate(data):
    s = int(r[ci["# Samples"]]); ins = int(r[ci["Instructions Executed"]])
    op = r[ci["Source"]].split()[0] if r[ci["Source"]].split() else ""
    if op.startswith("@"):
        op = r[ci["Source"]].split()[1]
    seg_s += s; seg_i += ins
    if op.startswith(("DFMA", "DMUL", "DADD")):
        seg_f += ins
    if op.startswith(("BAR", "EXIT")) or k == len(data) - 1:
        if seg_s > tot_s * 0.005 or seg_i > tot_i * 0.005:
            print("  sass %5d-%5d  samples %5.1f%%  instr %5.1f%%  fp64 instr %5.1f%%  (ends %s)" % (
                start, k, 100.0 * seg_s / tot_s, 100.0 * seg_i / tot_i, 100.0 * seg_f / tot_i, op))
        seg_s = seg_i = seg_f = 0
        start = k + 1
top = sorted(range(len(data)), key=lambda k: -int(data[k][ci["# Samples"]]))[:25]
print("top lines by samples:")
for k in top:
    r = data[k]
    print("  %5d %6.2f%%  exec %10s  thr/inst %5s  %s" % (k, 100.0 * int(r[ci["# Samples"]]) / tot_s, r[ci["Instructions Executed"]],
                                                      r[ci["Avg. Threads Executed"]], r[ci["Source"]].strip()[:90]))
