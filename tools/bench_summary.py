import json, sys
d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
r = d["roofline"]
print("ms/step %.2f  mol/s %.2f  step_frac %.3f  dom %s frac %.3f  e2e %.2f" % (d["ms_per_step"], d["value"], r["whole_step"]["frac"], r["kernel"], r["frac"], d["e2e"]["value"]))
print("  " + "  ".join("%s=%.2f" % (k, v["ms_per_step"]) for k, v in d["kernel_breakdown"].items()))
g = d["gin"]
print("gin graphs/s %.0f  ms %.3f  agg_frac %.3f  mlp_tf %.0f  e2e %.0f" % (g["value"], g["ms_per_forward"], g["roofline"]["frac"], g["mlp_gemms"]["tflops"] or 0, g["e2e"]["value"]))
print("  " + "  ".join("%s=%.3f" % (k, v["ms_per_forward"]) for k, v in g["kernel_breakdown"].items()))
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"]["value"], g.get("cpu_baseline", {}).get("value"))
print("clocks", d["clocks"])

if d.get("predictor"):
    q = d["predictor"]
    print("predictor graphs/s %.0f  ms %.2f  head_frac %s  e2e %.0f" % (q["value"], q["ms_per_batch"], q["roofline"]["frac"], q["e2e"]["value"]))
    print("  " + "  ".join("%s=%.3f" % (k, v["ms_per_batch"]) for k, v in q["kernel_breakdown"].items()))
