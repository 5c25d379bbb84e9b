"""Print the headline numbers and the per-kernel-class table of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.2f %s  ms/step %.2f  e2e %.2f  clocks %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]))
r = d["roofline"]
print("roofline: %s %.1f / %.1f %s  frac %.3f  share %.3f" % (r["bound"], r["achieved"], r["peak"], r["unit"], r["frac"], r.get("gemm_share_of_step", 0)))
for k, v in d.get("roofline_kernels", {}).items():
    print("  %-16s %8.1f us x %5.0f  frac %.3f  share %.4f" % (k, v["avg_us"], v["launches_per_step"], v["frac"], v["share_of_step"]))
for k, v in d.get("secondary", {}).items():
    print("  secondary %s: %s" % (k, {kk: vv for kk, vv in v.items() if kk in ("value", "unit", "ms_per_step")}))
