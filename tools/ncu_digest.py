"""Compact digest of an .ncu-rep (run here, no GPU): per captured kernel the bound-relevant metrics and the top warp-stall reasons."""
import csv, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[0], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
KEYS = [
    ("time us", "gpu__time_duration.sum", 1e-3), ("grid", "launch__grid_size", 1), ("regs", "launch__registers_per_thread", 1),
    ("warp-instr M", "sm__inst_executed.sum", 1e-6), ("issue-active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    ("warps-active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("fma pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    ("fmaheavy %", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("fma cycles %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("alu pipe %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    ("lsu wavefronts %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
    ("smem ld wavefronts M", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", 1e-6),
    ("smem st wavefronts M", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", 1e-6),
    ("smem ld conflicts M", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", 1e-6),
    ("smem st conflicts M", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", 1e-6),
    ("dram read MB", "dram__bytes_read.sum", 1e-6), ("dram write MB", "dram__bytes_write.sum", 1e-6),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("l2 %", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", 1),
]
for r in data:
    print("==", r[col["Kernel Name"]][:60], "id", r[col["ID"]])
    for label, key, sc in KEYS:
        if key in col and r[col[key]] not in ("", "n/a"):
            print(f"  {label:22s} {float(r[col[key]].replace(',', '')) * sc:12.2f}")
    stalls = []
    for n, i in col.items():
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a"):
            stalls.append((float(r[i].replace(",", "")), n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    if not stalls:
        for n, i in col.items():
            if "warp_issue_stalled" in n and n.endswith(".pct") and r[i] not in ("", "n/a"):
                stalls.append((float(r[i].replace(",", "")), n))
    for v, n in sorted(stalls, reverse=True)[:7]:
        print(f"  stall {n:40s} {v:8.3f}")
