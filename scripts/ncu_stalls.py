"""Per-kernel stall summary + top stalled SASS instructions of an .ncu-rep (developer tool): ncu_stalls.py <rep> [top]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
rx = re.compile(r"gpu__time_duration.sum|sm__cycles_elapsed.avg.per_second|sm__cycles_active.avg$|sm__cycles_elapsed.avg$|"
                r"smsp__issue_active.avg.pct|smsp__inst_executed.sum$|issue_stalled.*per_issue_active.ratio$|"
                r"sm__pipe_(alu|fma|fmaheavy|fmalite|tensor|xu).*cycles_active.avg.pct_of_peak_sustained_active|"
                r"dram__bytes_(read|write).sum$|smsp__warps_eligible.avg.per_cycle_active|launch__registers")
for r in rows[2:]:
    print("====", r[hdr.index("Kernel Name")][:60])
    for i, h in enumerate(hdr):
        if rx.search(h):
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            if v > 0.02:
                print(f"  {h.replace('smsp__average_warps_issue_stalled_', 'stall ').replace('_per_issue_active.ratio', '')} = {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name":
        name, hdr = rows[k][1], rows[k + 1]
        blk = []
        k += 2
        while k < len(rows) and not (rows[k] and rows[k][0] == "Kernel Name"):
            if len(rows[k]) == len(hdr):
                blk.append(rows[k])
            k += 1
        i_src, i_s = hdr.index("Source"), hdr.index("# Samples")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[i_s]) for r in blk)
        print("====", name[:60], "samples", tot)
        for idx, r in sorted(enumerate(blk), key=lambda x: -int(x[1][i_s]))[:top]:
            why = sorted(((int(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
            print(f"  {idx:5d} {r[i_src].strip()[:64]:64s} {100 * int(r[i_s]) / tot:5.1f}%  " + ", ".join(f"{h}={v}" for v, h in why if v))
    else:
        k += 1
