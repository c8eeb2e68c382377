"""ncu --page raw --csv -> one line per captured launch with the metrics the round summary quotes."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dsmem")]
cols = [(c, n) for c, n in cols if c in ix]
print("kernel".ljust(44), *[n.rjust(11) for _, n in cols])
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
    name = re.sub(r"ua::<unnamed>::|<unnamed>::", "", name)[:44]
    vals = []
    for c, n in cols:
        v, u = r[ix[c]], units[ix[c]]
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.4g}" + ({"Mbyte": "MB", "Kbyte": "KB", "Gbyte": "GB", "byte": "B", "us": "us", "ms": "ms", "ns": "ns", "%": "", "msecond": "ms", "usecond": "us"}.get(u, ""))
        except ValueError:
            pass
        vals.append(v.rjust(11))
    print(name.ljust(44), *vals)
