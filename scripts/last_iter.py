"""Per-launch device times of the LAST iteration in an ncu launch-list CSV (gpu__time_duration.sum): python scripts/last_iter.py file.csv [marker=round_kernel]"""
import csv
import sys

for f in sys.argv[1:]:
    if not f.endswith(".csv"):
        continue
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = rows[0]
    iN, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    out = []
    for r in rows[1:]:
        v = float(r[iV].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r[iU], v)
        out.append((r[iN][:60], v))
    idx = [i for i, (n, v) in enumerate(out) if "round_kernel" in n]
    start = idx[-1] if idx else 0
    print(f, "last iteration:")
    tot = 0
    for n, v in out[start:]:
        print(f"   {v:9.1f} us  {n}")
        tot += v
    print("   total %.1f us in %d launches" % (tot, len(out) - start))
