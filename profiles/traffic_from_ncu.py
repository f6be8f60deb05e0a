"""Turn the one-pass ncu CSVs of profiles/r2_traffic.sh into per-config steady-state DRAM bytes per launch.

    python profiles/traffic_from_ncu.py <dir with traffic_<key>.csv> [--write]

Prints a markdown table; with --write also updates profiles/traffic.json (keys `<label>_T<T>_N<N>`, the ones
bench.py looks up) with the mean over the captured launches."""
import csv
import glob
import json
import os
import sys

ALGO = {1: 8804, 10: 40772}
d = sys.argv[1]
rows = []
out = {}
for path in sorted(glob.glob(os.path.join(d, "traffic_*.csv"))):
    key = os.path.basename(path)[len("traffic_"):-len(".csv")]
    per = {}
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = r.get("Metric Unit", "")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1,
                 "msecond": 1e6, "ms": 1e6}.get(unit, 1)
        per.setdefault(r["ID"], {"kernel": r["Kernel Name"]})[r["Metric Name"]] = v * scale
    if not per:
        continue
    n = len(per)
    rd = sum(p.get("dram__bytes_read.sum", 0) for p in per.values()) / n
    wr = sum(p.get("dram__bytes_write.sum", 0) for p in per.values()) / n
    l2 = sum(p.get("lts__t_bytes.sum", 0) for p in per.values()) / n
    ns = sum(p.get("gpu__time_duration.sum", 0) for p in per.values()) / n
    kern = next(iter(per.values()))["kernel"]
    N = int(key.split("_N")[1])
    T = int(key.split("_T")[1].split("_")[0])
    algo = (10052 if key.startswith("config4") else ALGO[T]) * N
    out[key] = int(rd + wr)
    rows.append((key, kern[:40], n, rd / 1e6, wr / 1e6, (rd + wr) / 1e6, algo / 1e6, (rd + wr) / algo, l2 / 1e6, ns / 1e3))
print("| config | kernel | launches | DRAM read MB | DRAM write MB | DRAM total MB | algorithmic MB | DRAM / algorithmic | L2 traffic MB | us under ncu (serialised) |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.2f} | {r[4]:.2f} | {r[5]:.2f} | {r[6]:.2f} | {r[7]:.3f} | {r[8]:.1f} | {r[9]:.2f} |")
if "--write" in sys.argv:
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    out["source"] = ("profiles/r2_traffic.sh: ncu one-pass capture (dram__bytes_read.sum + dram__bytes_write.sum, --cache-control none, "
                     "--clock-control none), mean over launches 40..47 of a 60-launch run over a ring of buffer sets larger than L2 "
                     "(steady state: two laps of the ring before the first captured launch)")
    json.dump(out, open(p, "w"), indent=1)
