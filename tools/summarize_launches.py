"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel table: launches, total us, share of the capture."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr_i + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").strip()
    name = re.sub(r"ua::<unnamed>::|ua::\(anonymous namespace\)::", "ua::", name)
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] in ("ns", "nsecond") else (v * 1e3 if r[mu] in ("ms", "msecond") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':80s} {'launches':>8s} {'total_us':>12s} {'mean_us':>10s} {'share':>7s}")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:80]:80s} {n:8d} {t:12.1f} {t / n:10.2f} {100 * t / tot:6.1f}%")
print(f"{'TOTAL':80s} {sum(a[0] for a in agg.values()):8d} {tot:12.1f}")
