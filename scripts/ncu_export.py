"""Summarise an `ncu --set full` report of the pass kernel into profiles/ (run in the build container):
    python scripts/ncu_export.py gpurun_out/<call>/prof_ts.ncu-rep r2_rsvd_ts [cells genes n_gpus]
writes profiles/<tag>_metrics.csv (selected raw metrics per captured launch), profiles/<tag>_stalls.txt (warp-stall
sampling summary, opcode mix per matrix entry, hottest instructions) and profiles/r2_rsvd_tc_traffic.json (DRAM bytes per
launch, read by bench.py's roofline block)."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
cells, genes, n_gpus = (int(a) for a in (sys.argv[3:6] or ["50000", "20000", "1"]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
want = [w for w in want if w in idx]
with open(f"profiles/{tag}_metrics.csv", "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(want)
    w.writerow([units[idx[k]] for k in want])
    for r in rows[2:]:
        w.writerow([r[idx[k]] for k in want])


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


traffic = [to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) +
           to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]) for r in rows[2:]]
json.dump({"workload": {"cells": cells, "genes": genes, "n_gpus": n_gpus}, "kernel": "rsvd_apply_ts_kernel",
           "launches_captured": len(traffic), "bytes_per_launch": traffic,
           "mean_bytes_per_launch": int(sum(traffic) / len(traffic)), "source": rep},
          open("profiles/r2_rsvd_tc_traffic.json", "w"), indent=1)

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
out, seen = [], set()
entries = float(cells) * genes
for k in kern:
    if k["name"] in seen:
        continue
    seen.add(k["name"])
    h = {n: i for i, n in enumerate(k["hdr"])}
    R = k["rows"]
    tot = sum(int(r[h["# Samples"]] or 0) for r in R)
    ie = sum(int(r[h["Instructions Executed"]] or 0) for r in R)
    out.append(f"== {k['name']}\n   warp instructions {ie} = {ie * 32 / entries:.2f} lane-instructions per matrix entry; {tot} stall samples")
    st = collections.Counter()
    for r in R:
        for c in k["hdr"]:
            if c.startswith("stall_") and "Not Issued" not in c:
                st[c] += int(r[h[c]] or 0)
    out.append("   stall samples: " + ", ".join(f"{c[6:]} {v} ({100 * v / max(tot, 1):.0f} %)" for c, v in st.most_common(8)))
    ops = collections.Counter()
    for r in R:
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[h["Source"]])
        ops[m.group(1).split(".")[0] if m else "?"] += int(r[h["Instructions Executed"]] or 0)
    out.append("   per entry: " + " ".join(f"{op}:{c * 32 / entries:.2f}" for op, c in ops.most_common(26)))
    for i in sorted(sorted(range(len(R)), key=lambda i: -int(R[i][h["# Samples"]] or 0))[:8]):
        out.append(f"   hot: {R[i][h['# Samples']]:>6} samples  {R[i][h['Source']].strip()[:80]}")
open(f"profiles/{tag}_stalls.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
