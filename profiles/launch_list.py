"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` log into launches / total / share per kernel."""
import csv, sys, re, collections
rows = collections.OrderedDict()
with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("(anonymous namespace)::", "").strip()
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
    n, t = rows.get(name, (0, 0.0))
    rows[name] = (n + 1, t + ms)
tot = sum(t for _, t in rows.values())
print(f"{'kernel':40s} {'launches':>8s} {'total ms':>10s} {'ms/launch':>10s} {'share':>7s}")
for name, (n, t) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:40]:40s} {n:8d} {t:10.3f} {t / n:10.3f} {100 * t / tot:6.1f}%")
