#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` into a markdown table (one row per profiled launch)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"), ("lts__t_bytes.sum", "L2 bytes"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("smsp__inst_executed.sum", "warp inst")]
print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    short = re.sub(r"^void (xtb::)?(tc::)?", "", name)
    short = re.sub(r"\(int\)|\(bool\)|xtb::|tc::", "", short)
    short = re.sub(r"\(T\d.*$", "", short)[:110]
    vals = []
    for k, _ in cols:
        if k in idx:
            v = r[idx[k]]
            try:
                f = float(v.replace(",", ""))
                v = ("%.3g" % f) + " " + units[idx[k]]
            except ValueError:
                pass
            vals.append(v)
        else:
            vals.append("-")
    print("| `%s` | " % short + " | ".join(vals) + " |")
