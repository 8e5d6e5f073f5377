#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + per-phase attribution) into a small text file for profiles/."""
import csv, subprocess, sys, os, tempfile
rep, out, label = sys.argv[1], sys.argv[2], sys.argv[3]
sass, pat = (sys.argv[4], sys.argv[5]) if len(sys.argv) > 5 else (None, None)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
with open(out, "w") as f:
    f.write("# %s\n# source: %s (ncu --set full --clock-control none; per-launch values, cold-ish cache, serialised)\n" % (label, os.path.basename(rep)))
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        f.write("\nkernel: %s\n" % name[:160])
        for k in keys:
            if k in hdr:
                f.write("  %-72s %16s %s\n" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        f.write("  stalls per issued instruction:\n")
        for h in sorted(stall, key=lambda h: -float(r[hdr.index(h)] or 0))[:8]:
            f.write("    %-40s %8.3f\n" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[hdr.index(h)])))
    if sass:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as t:
            t.write(src)
        ph = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_by_phase.py"), sass, pat, t.name],
                            capture_output=True, text=True).stdout
        f.write("\nper-phase attribution (outermost call-site line in ram_fused.cu; warp-instructions, stall samples):\n")
        for ln in ph.splitlines():
            parts = ln.split()
            if ln.startswith("total") or (len(parts) > 2 and parts[1] == "inst" and float(parts[2].rstrip("%")) >= 0.3):
                f.write("  " + ln[:220] + "\n")
print("wrote", out)
