#!/usr/bin/env python
"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --no-graph` into the per-launch
table of ONE forward (the last complete one in the capture) for profiles/.  usage: launch_summary.py in.csv out.md title"""
import csv, sys

NAMES = ["conv1_7x7_s2 + bn + relu + pool1 (frames in)", "conv2_3x3_reduce", "conv2_3x3", "pool2",
         "3a: 1x1 + 3x3_reduce + double_3x3_reduce + pool_proj (one GEMM)", "3a_3x3", "3a_double_3x3_1", "3a_double_3x3_2",
         "3a_pool (after pool_proj: +bias+BN+ReLU)",
         "3b: 1x1 + 3x3_reduce + double_3x3_reduce + pool_proj (one GEMM)", "3b_3x3", "3b_double_3x3_1", "3b_double_3x3_2",
         "3b_pool (after pool_proj: +bias+BN+ReLU)", "3c_double_3x3_reduce", "3c_double_3x3_1",
         "res3a_2n", "res3b_1", "res3b_2 (+res3a)", "res4a_1", "res4a_2", "res4a_down (+res4a_2)", "res4b_1", "res4b_2 (+res4a)",
         "res5a_1", "res5a_2", "res5a_down (+res5a_2)", "res5b_1", "res5b_2 (+res5a)", "global_pool", "fc8"]

def main(src, dst, title):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            name = r["Kernel Name"].split("(")[0].replace("eco::<unnamed>::", "").replace("void ", "")
            rows.append((name, float(r["Metric Value"].replace(",", "")) / 1000.0, r["Grid Size"], r["Block Size"]))
    starts = [i for i, r in enumerate(rows) if "stem_rows" in r[0]]
    assert len(starts) >= 2, "need at least two forwards in the capture"
    # the capture also holds the chunked e2e forwards (sub-batches on sub-nets): keep to the full-batch forwards, i.e. the
    # ones whose stem launch has the largest grid, and take the last complete one of them
    def grid0(i):
        return int(rows[i][2].strip("()").split(",")[0])
    gmax = max(grid0(i) for i in starts)
    full = [k for k, i in enumerate(starts[:-1]) if grid0(i) == gmax and grid0(starts[k + 1]) == gmax]
    assert full, "no complete full-batch forward in the capture"
    fw = rows[starts[full[-1]]:starts[full[-1] + 1]]
    tot = sum(r[1] for r in fw)
    agg = {}
    for n, t, g, b in fw:
        a = agg.setdefault(n, [0.0, 0]); a[0] += t; a[1] += 1
    with open(dst, "w") as f:
        f.write("# %s\n\n" % title)
        f.write("per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.  "
                "%d launches, sum %.1f us.\n\n| kernel | launches | us | share |\n|---|---|---|---|\n" % (len(fw), tot))
        for n, (t, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
            f.write("| %s | %d | %.1f | %.1f%% |\n" % (n[:70], c, t, 100 * t / tot))
        f.write("\n| # | layer(s) | kernel | grid | block | us |\n|---|---|---|---|---|---|\n")
        for i, (n, t, g, b) in enumerate(fw):
            f.write("| %d | %s | %s | %s | %s | %.1f |\n" % (i, NAMES[i] if i < len(NAMES) else "", n[:44], g, b, t))
    print(open(dst).read()[:900])

if __name__ == "__main__":
    main(*sys.argv[1:4])
