"""Per-kernel-class DRAM traffic of one training step from an ncu capture:
     ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size \
         --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/profile_step.py
   python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/r02_traffic.json
Writes {kind: {"dram_bytes_per_launch": avg, "launches": n, "source": ...}} with the kinds bench.py's roofline_kernels uses."""
import collections
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
with open(src) as f:
    lines = [l for l in f if not l.startswith("==")]
per = collections.defaultdict(dict)
for r in csv.DictReader(lines):
    key = (r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]))
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
    per[key][r["Metric Name"]] = v * scale


def kind_of(name, m):
    grid = m.get("launch__grid_size", 0)
    if "gemm_tcgen05" in name:
        return "gemm"
    if "attn_fwd" in name:
        return "attn_self_fwd" if grid <= 512 else "attn_cross_fwd"
    if "attn_bwd_dq" in name:
        return "attn_self_bwd" if grid <= 512 else "attn_cross_bwd"
    if "attn_bwd_dkv" in name:
        return "attn_self_bwd" if grid <= 1024 else "attn_cross_bwd"
    if "add_ln_fwd" in name:
        return "add_ln_fwd"
    if "add_ln_bwd" in name:
        return "add_ln_bwd"
    if "embed_ln_fwd" in name:
        return "embed_ln_fwd"
    if "ce_fwd_bwd" in name:
        return "ce_bwd" if m.get("dram__bytes_write.sum", 0) > 1e8 else "ce_fwd"
    if "colsum" in name:
        return "colsum"
    if "gate_" in name:
        return "gate"
    return None


agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (_, name), m in per.items():
    k = kind_of(name, m)
    if k is None:
        continue
    a = agg[k]
    a[0] += 1
    a[1] += m.get("dram__bytes_read.sum", 0.0)
    a[2] += m.get("dram__bytes_write.sum", 0.0)
    a[3] += m.get("gpu__time_duration.sum", 0.0)
out = {}
for k, (n, rd, wr, ns) in sorted(agg.items()):
    # the two backward attention kernels form ONE bench "launch" (one C-ABI call)
    calls = n / 2 if k.endswith("_bwd") and k.startswith("attn") else n
    out[k] = {"dram_bytes_per_launch": (rd + wr) / calls, "dram_read_bytes_per_launch": rd / calls, "dram_write_bytes_per_launch": wr / calls,
              "launches": calls, "ncu_us_per_launch": ns / calls / 1e3,
              "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, one training step at the config-2 shape (%s)" % src}
json.dump(out, open(dst, "w"), indent=1)
for k, v in out.items():
    print("%-16s %5d launches  %9.1f MB/launch  %8.1f us" % (k, v["launches"], v["dram_bytes_per_launch"] / 1e6, v["ncu_us_per_launch"]))
