"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/launch_table.py in.csv out.csv [regex]"""
import collections, csv, re, sys
src, dst = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
        continue
    name = r[ki].split("(")[0]
    if pat and not pat.search(name):
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(r[ui], 1.0)
    tot[name] += v; cnt[name] += 1
T = sum(tot.values())
with open(dst, "w") as f:
    f.write("kernel,launches,total_us,share_pct,avg_us\n")
    for k, v in tot.most_common():
        f.write(f"{k},{cnt[k]},{v:.1f},{100 * v / T:.2f},{v / cnt[k]:.2f}\n")
print("wrote", dst, "kernels", len(tot), "total_us", round(T, 1))
