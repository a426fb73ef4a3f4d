"""Per-launch device times from an `ncu --metrics gpu__time_duration.sum --csv` log (second half = last build)."""
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else v * 1e3 if row["Metric Unit"] == "ms" else v
        rows.append((int(row["ID"]), name, v, row["Grid Size"]))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
start = len(rows) - len(rows) // reps if reps > 1 else 0
tot = 0.0
for id_, name, us, grid in rows[start:]:
    if "check" in name:
        continue
    print(f"{id_:4d} {us:10.1f} us  {name[:72]:72s} {grid}")
    tot += us
print(f"total {tot:.1f} us over {len(rows) - start} launches")
