"""profiles/radix_pass_traffic.json from an ncu report that captured every radix_pass launch of ONE build
(ncu --set full -k "regex:radix_pass|ds_local_kernel" -c <launches per build>):
DRAM bytes (read + write) per launch, averaged over the launches like bench.py's `achieved`.
    python tools/traffic_from_ncu.py report.ncu-rep out.json"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def val(d, name):
    v = float(d[ix[name]].replace(",", ""))
    u = units[ix[name]]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


launches, text = [], []
for d in data:
    name = d[ix["Kernel Name"]]
    if "radix_pass" not in name and "ds_local_kernel" not in name:
        continue  # the launches of the ESA_K_RADIX_PASS family (ds_local_kernel is the in-tile sort of a round)
    entry = {"grid": d[ix["Grid Size"]], "kernel": name[:80],
             "dram_bytes": val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum"),
             "time_us": float(d[ix["gpu__time_duration.sum"]].replace(",", "")) *
             {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1)}
    # the text-keyed first pass is a family of its own (ESA_K_RADIX_PASS_TEXT), like in bench.py's roofline
    (text if "radix_pass_dna" in name else launches).append(entry)
tot = sum(l["dram_bytes"] for l in launches)
tot_text = sum(l["dram_bytes"] for l in text)
json.dump({"source": rep, "launches": len(launches), "dram_bytes_per_launch": tot / max(1, len(launches)),
           "dram_bytes_total": tot, "per_launch": launches,
           "text_pass": {"launches": len(text), "dram_bytes_per_launch": tot_text / max(1, len(text)),
                         "per_launch": text}}, open(out, "w"), indent=1)
print(len(launches), "launches,", tot / max(1, len(launches)) / 1e6, "MB per launch;", len(text), "text-keyed,",
      tot_text / max(1, len(text)) / 1e6, "MB per launch")
