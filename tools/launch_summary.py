"""Per-kernel shares from an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`).
  python tools/launch_summary.py X.csv > summary.txt"""
import csv, re, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot = {}
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1e-6)  # -> ms
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:100]
    n, t = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, t + v)
total = sum(t for _, t in tot.values())
print(f"{'kernel':<100} {'launches':>9} {'ms total':>10} {'share':>7}")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:<100} {n:>9} {t:>10.3f} {100 * t / total:>6.2f}%")
