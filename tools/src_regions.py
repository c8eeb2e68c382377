"""ncu --page source --csv (SASS view) -> instructions per warp and stall samples, by opcode and by address range:
python tools/src_regions.py file.csv [n_ranges]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {'name': r[1], 'rows': []}; ks.append(cur)
    elif r and r[0] == "Address": cur['hdr'] = r
    elif cur is not None and r: cur['rows'].append(r)
for k in ks:
    h = k['hdr']; ii = h.index("Instructions Executed"); si = h.index("Warp Stall Sampling (All Samples)")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    R = k['rows']; w = float(R[0][ii]) or 1.0
    tot = sum(float(r[ii]) for r in R); ts = sum(float(r[si]) for r in R) or 1
    print(k['name'][:100]); print(f"  SASS lines {len(R)}, warps {w:.0f}, instructions per warp {tot / w:.1f}, samples {ts:.0f}")
    op = collections.Counter(); st = collections.Counter()
    for r in R:
        o = [x for x in r[1].split() if not x.startswith('@')][0].split('.')[0]
        op[o] += float(r[ii]) / w; st[o] += float(r[si])
    print("  by opcode: " + "  ".join(f"{o} {c:.0f} ({100 * st[o] / ts:.0f}%)" for o, c in op.most_common(22)))
    n = len(R); step = (n + nr - 1) // nr
    for a in range(0, n, step):
        seg = R[a:a + step]
        ins = sum(float(r[ii]) for r in seg) / w; sm = sum(float(r[si]) for r in seg)
        stalls = collections.Counter()
        for r in seg:
            for c in stall_cols:
                try: stalls[h[c]] += float(r[c])
                except ValueError: pass
        top = ", ".join(f"{n_[6:]} {100 * v / max(sm, 1):.0f}%" for n_, v in stalls.most_common(3))
        print(f"  lines {a:5d}-{a + len(seg) - 1:5d}: {ins:7.1f} inst/warp  {100 * sm / ts:5.1f}% samples  [{top}]  first: {seg[0][1].strip()[:50]}")
