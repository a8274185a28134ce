#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics and the hottest SASS lines by stall samples.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [n_hot]"""
import csv, subprocess, sys
rep = sys.argv[1]; nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "smsp__thread_inst_executed.sum"]
for vals in rows[2:]:
    print("kernel:", vals[hdr.index("Kernel Name")][:100])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f"  {w:70s} {vals[i]:>18s} {units[i]}")
    # warps stalled per issue-active cycle, by reason
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or \
           h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio"):
            try: st.append((float(vals[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "").replace(".ratio", "")))
            except ValueError: pass
    st.sort(reverse=True)
    if st: print("  stalled warps per issue-active cycle: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if h:
    hdr = rows[h[0]]; data = [r for r in rows[h[0] + 1:(h[1] if len(h) > 1 else len(rows))] if len(r) == len(hdr)]
    si = hdr.index("# Samples"); so = hdr.index("Source"); ie = hdr.index("Instructions Executed")
    tot = sum(int(r[si]) for r in data if len(r) > si and r[si].isdigit())
    toti = sum(int(r[ie]) for r in data if len(r) > ie and r[ie].isdigit())
    print(f"  total stall samples {tot}, warp instructions executed {toti}, SASS lines {len(data)}")
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][si]) if data[i][si].isdigit() else 0)[:nhot]
    for i in sorted(idx):
        print(f"  [{i:5d}] samples {data[i][si]:>6s} exec {data[i][ie]:>10s}  {data[i][so].strip()[:100]}")
