#!/usr/bin/env python
"""SURVEY.md 8(d) config #2 asks for B in {1, 8, 64} and the best B: run bench.py (device-timed `value`, e2e keys) for a list of
batch sizes at the current kernels and write a markdown / json table.
Usage: python tools/batch_sweep.py [--batches 1 8 16 32 64] [--model lite] [--out gpurun_out/batch_sweep]"""
import argparse, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--batches", type=int, nargs="+", default=[1, 8, 16, 32, 64])
ap.add_argument("--model", default="lite")
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "batch_sweep"))
a = ap.parse_args()
rows = []
for b in a.batches:
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--batch", str(b), "--model", a.model, "--steps", str(a.steps),
                        "--warmup", "5", "--no-cpu-baseline"], capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        rows.append({"batch": b, "error": r.stderr[-500:]})
        continue
    d = json.loads(line[-1])
    rows.append({"batch": b, "value": d["value"], "ms_per_step": d["ms_per_step"], "e2e": d["e2e"]["value"],
                 "e2e_pipelined": d["e2e"]["pipelined"]["value"], "e2e_u8": d["e2e_u8"]["value"],
                 "roofline_frac": d["roofline"]["frac"], "whole_step_tflops": d["roofline"]["whole_step_tflops"],
                 "sm_mhz": d["clocks"]["sm_mhz"], "reasons": d["clocks"]["reasons"]})
json.dump({"model": a.model, "rows": rows}, open(a.out + ".json", "w"), indent=1)
best = max((r for r in rows if "value" in r), key=lambda r: r["value"])
with open(a.out + ".md", "w") as f:
    f.write("# ECO-%s N=16 forward, batch sweep on 1 x B200 (bench.py, device-timed value; e2e = host buffers in, logits out)\n\n" % a.model)
    f.write("| videos/step | value videos/s | ms/step | conv roofline frac | whole-step TFLOP/s | e2e blocking fp32 | e2e pipelined fp32 | e2e pipelined uint8 |\n|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        if "value" in r:
            f.write("| %d | %.0f | %.3f | %.3f | %.0f | %.0f | %.0f | %.0f |\n" % (r["batch"], r["value"], r["ms_per_step"], r["roofline_frac"],
                                                                            r["whole_step_tflops"], r["e2e"], r["e2e_pipelined"], r["e2e_u8"]))
        else:
            f.write("| %d | failed: %s |\n" % (r["batch"], r["error"].replace("\n", " ")[-200:]))
    f.write("\nbest batch: %d videos/step (%.0f videos/s); latency at B=1: %.3f ms per clip\n" % (
        best["batch"], best["value"], next((r["ms_per_step"] for r in rows if r.get("batch") == 1 and "value" in r), float("nan"))))
print(open(a.out + ".md").read())
