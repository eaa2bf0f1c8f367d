#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`) by
kernel name: launches, time, share of the step and — when the DRAM metrics are present — bytes moved and GB/s.
With `--json FILE` also writes the per-step DRAM traffic of the tcgen05 GEMM class (bench.py's roofline.traffic)."""
import csv
import json
import re
import sys
from collections import defaultdict

args = [a for a in sys.argv[1:] if not a.startswith("--")]
path = args[0]
json_out = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
per = defaultdict(dict)          # launch id -> {name, us, rd, wr}
for r in csv.DictReader(lines):
    d = per[r["ID"]]
    d["name"] = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    m = r.get("Metric Name")
    if m == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d["rd" if m.endswith("read.sum") else "wr"] = v * scale
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for d in per.values():
    a = agg[d["name"]]
    a[0] += 1; a[1] += d.get("us", 0.0); a[2] += d.get("rd", 0.0); a[3] += d.get("wr", 0.0)
total = sum(a[1] for a in agg.values())
have_dram = any(a[2] + a[3] > 0 for a in agg.values())
print(f"total {total/1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches"
      + (f", DRAM {sum(a[2] for a in agg.values())/1e9:.2f} GB read + {sum(a[3] for a in agg.values())/1e9:.2f} GB written" if have_dram else ""))
for name, (n, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    extra = f"  {(rd+wr)/1e6:9.1f} MB  {(rd+wr)/max(us,1e-9)/1e3:6.0f} GB/s" if have_dram else ""
    print(f"{us/1e3:9.3f} ms {100*us/total:5.1f}%  n={n:4d}  avg={us/n:9.1f} us{extra}  {name[:100]}")
if json_out:
    g = [a for k, a in agg.items() if "gemm_tc_kernel" in k]
    out = {"source": path, "kernel": "gemm_tc_kernel (all instantiations)", "launches_per_step": sum(a[0] for a in g),
           "dram_bytes_per_step": sum(a[2] + a[3] for a in g), "ms_per_step_under_ncu": sum(a[1] for a in g) / 1e3,
           "step_dram_bytes_all_kernels": sum(a[2] + a[3] for a in agg.values())}
    json.dump(out, open(json_out, "w"), indent=1)
