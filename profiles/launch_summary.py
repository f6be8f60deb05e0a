"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python profiles/launch_summary.py gpurun_out/launches.csv [out.md] [--tail N]

With --tail N only the last N launches are aggregated (the timed graph replays of bench.py come
last, after workload generation and warm-up)."""
import collections
import csv
import sys

path = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else sys.stdout
tail = int(sys.argv[sys.argv.index("--tail") + 1]) if "--tail" in sys.argv else None
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"].split("(")[0][:90], ns))
if tail:
    rows = rows[-tail:]
agg = collections.defaultdict(lambda: [0, 0.0])
for name, ns in rows:
    agg[name][0] += 1
    agg[name][1] += ns
total = sum(v[1] for v in agg.values())
print(f"# {path}: {len(rows)} launches, {total / 1e3:.1f} us of kernel time"
      + (f" (last {tail} launches)" if tail else "") + "\n", file=out)
print("| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|", file=out)
for name, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:15]:
    print(f"| `{name}` | {cnt} | {ns / 1e3:.1f} | {100 * ns / total:.1f} % | {ns / cnt / 1e3:.2f} |", file=out)
