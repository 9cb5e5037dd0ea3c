#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box): one row per launch with the metrics the roofline needs."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units = r[0], r[1]
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "us"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_sector_hit_rate.pct", "L2hit%"), ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
        ("lts__t_bytes.sum", "L2bytes"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("sm__cycles_elapsed.avg.per_second", "clk"), ("launch__registers_per_thread", "regs"),
        ("smsp__inst_executed.sum", "inst")]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
print(" | ".join(n for _, n in idx))
for row in r[2:]:
    out = []
    for i, n in idx:
        v = row[i]
        if n == "kernel":
            v = v.split("(")[0].split("::")[-1][:28]
        elif units[i] in ("byte", "Kbyte", "Mbyte", "Gbyte"):
            v = "%s%s" % (v, units[i][0] if units[i] != "byte" else "B")
        out.append(v)
    print(" | ".join(out))
