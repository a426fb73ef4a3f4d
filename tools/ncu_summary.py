"""Summarise an .ncu-rep: per kernel the headline metrics, and for one kernel the opcode mix / stall reasons.
    python tools/ncu_summary.py report.ncu-rep [kernel-regex-for-source-page]"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "inst"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%")]
seen = set()
for d in data:
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", d[ix["Kernel Name"]])
    name = re.sub(r"\(.*", "", name)
    key = (name, d[ix["Grid Size"]])
    if key in seen:
        continue
    seen.add(key)
    parts = []
    for m, short in want:
        if m in ix:
            v = d[ix[m]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            parts.append(f"{short}={v}{units[ix[m]] if short in ('time', 'dram_rd', 'dram_wr') else ''}")
    print(f"{name[:60]:60s} grid={d[ix['Grid Size']]:16s} " + " ".join(parts))

if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2],
                          "--launch-count", "1", "--launch-skip", sys.argv[3] if len(sys.argv) > 3 else "0"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    d = []
    for r in rows[2:]:  # the SASS view; a report with imported sources repeats the header for the next view
        if len(r) < len(h) or not r[h.index("# Samples")].isdigit():
            break
        d.append(r)
    jx = {x: i for i, x in enumerate(h)}
    tot = sum(int(r[jx["# Samples"]]) for r in d)
    toti = sum(int(r[jx["Instructions Executed"]]) for r in d)
    print("\nsource page:", rows[0][1][:100], "samples", tot, "warp-instructions", toti)
    c, s = Counter(), Counter()
    for r in d:
        t = r[jx["Source"]].split()
        op = t[1] if t[0].startswith("@") else t[0]
        op = op.split(".")[0]
        c[op] += int(r[jx["Instructions Executed"]])
        s[op] += int(r[jx["# Samples"]])
    for op, n in c.most_common(18):
        print(f"  {op:10s} inst {100 * n / toti:5.1f}%  samples {100 * s[op] / tot:5.1f}%")
    st = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    agg = sorted(((sum(int(r[jx[x]]) for r in d), x) for x in st), reverse=True)[:7]
    print("  stalls:", ", ".join(f"{x[6:]} {100 * v / tot:.0f}%" for v, x in agg))
    top = sorted(d, key=lambda r: -int(r[jx["# Samples"]]))[:12]
    for r in top:
        print(f"   {100 * int(r[jx['# Samples']]) / tot:5.1f}%  {r[jx['Source']][:90]}")
