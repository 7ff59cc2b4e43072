"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised: compare SHARES)."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    try: v = float(row["Metric Value"].replace(",", ""))
    except ValueError: continue
    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("%-58s %6s %12s %7s" % ("kernel", "n", "total ms", "share"))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("%-58s %6d %12.3f %6.1f%%" % (k[:58], n, t / 1e6, 100 * t / tot))
print("%-58s %6s %12.3f" % ("TOTAL", "", tot / 1e6))
