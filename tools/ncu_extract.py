#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (ncu --set full) as a small text table for profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
idx = [(h, hdr.index(h)) for h in want if h in hdr]
for r in rows[2:]:
    print("-" * 100)
    for h, i in idx:
        print(f"{h:72s} {units[i]:14s} {r[i][:120]}")
    try:
        t = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
        tu = units[hdr.index("gpu__time_duration.sum")]
        t_s = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(tu, 1e-9)
        def tobytes(name):
            v = float(r[hdr.index(name)].replace(",", "")); u = units[hdr.index(name)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        b = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
        print(f"{'derived: DRAM traffic (read+write)':72s} {'MB':14s} {b/1e6:.1f}")
        print(f"{'derived: achieved DRAM bandwidth':72s} {'GB/s':14s} {b/t_s/1e9:.0f}   ({100*b/t_s/6527.1e9:.1f}% of measured 6527 GB/s)")
    except Exception as e:
        print("derived metrics unavailable:", e)
