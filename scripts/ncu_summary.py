"""Summaries of ncu output (developer tool): `launches <csv>` aggregates a gpu__time_duration launch list;
`raw <ncu-rep> [metric-regex]` prints selected raw metrics per captured launch."""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3}[row["Metric Unit"]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    print(f"total {tot / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{t / 1e3:10.3f} ms {n:6d} x {t / n:9.1f} us {100 * t / tot:5.1f}%  {k[:100]}")


DEFAULT = ("gpu__time_duration.sum|dram__bytes_read.sum$|dram__bytes_write.sum$|gpu__dram_throughput.avg.pct|"
           "sm__throughput.avg.pct|sm__warps_active.avg.pct|launch__registers|launch__grid_size|launch__occupancy_limit|"
           "smsp__issue_active.avg.pct|sm__pipe_fma_cycles_active.avg.pct|sm__pipe_alu_cycles_active.avg.pct|"
           "sm__pipe_tensor.*cycles_active.avg.pct|sm__inst_executed_pipe_xu.sum$|smsp__inst_executed.sum$|"
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|lts__t_bytes.sum$|sm__pipe_xu_cycles_active.avg.pct|"
           "smsp__average_warp.*issue_stalled.*ratio$|achieved_occupancy|sm__pipe_fmaheavy|launch__waves")


def raw(path, pattern=DEFAULT):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    rx = re.compile(pattern)
    for row in rows[2:]:
        print("====", row[hdr.index("Kernel Name")][:90])
        for i, h in enumerate(hdr):
            if rx.search(h):
                print(f"  {h} = {row[i]} {units[i]}")




def opcodes(path, kernel_regex=".*", top=40):
    """Executed warp-instructions by SASS opcode for the first captured kernel whose name matches."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rx = re.compile(kernel_regex)
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and row:
            cur["rows"].append(row)
    for b in blocks:
        if not rx.search(b["name"]):
            continue
        h = b["hdr"]
        i_src, i_ex, i_smp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
        agg, tot, smp_tot = collections.Counter(), 0, 0
        smp = collections.Counter()
        for r in b["rows"]:
            toks = r[i_src].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            op = ".".join(op.split(".")[:2])
            n = int(r[i_ex] or 0)
            agg[op] += n
            tot += n
            smp[op] += int(r[i_smp] or 0)
            smp_tot += int(r[i_smp] or 0)
        print("====", b["name"][:80], "total warp-instr", tot)
        for op, n in agg.most_common(int(top)):
            print(f"  {op:28s} {n:12d} {100 * n / tot:5.1f}%   samples {100 * smp[op] / max(smp_tot, 1):5.1f}%")
        break

if __name__ == "__main__":
    {"launches": launches, "raw": raw, "opcodes": opcodes}[sys.argv[1]](*sys.argv[2:])
