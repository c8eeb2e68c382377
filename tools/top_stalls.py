"""ncu --page source --csv -> the source lines with the most warp-stall samples (needs -lineinfo + --import-source on)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
while rows and "Source" not in rows[0]:      # a "Kernel Name" line precedes the header
    rows.pop(0)
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
samp = "Warp Stall Sampling (All Samples)"
src = next((h for h in hdr if h in ("Source", "source")), hdr[1])
inst = "Instructions Executed"
print("columns:", samp, "|", inst, file=sys.stderr)
def num(v):
    try: return float(v.replace(",", ""))
    except Exception: return 0.0
data = [(num(r[ix[samp]]), num(r[ix[inst]]) if inst else 0, r[ix[src]].strip()[:150], r[0]) for r in rows[1:] if len(r) > ix[samp] and r[0] != "Address" and not r[0].startswith("Kernel")]
tot = sum(d[0] for d in data) or 1
toti = sum(d[1] for d in data) or 1
print(f"total samples {tot:.0f}, instructions {toti:.0f}")
for s, i, t, ln in sorted(data, reverse=True)[:n]:
    print(f"{100*s/tot:5.1f}% samp {100*i/toti:5.1f}% inst  L{ln}: {t}")
