"""Markdown summary of `ncu --set full` reports (run where ncu is installed; the .ncu-rep files come back from the
GPU box under gpurun_out/).  usage: python scripts/ncu_summary.py a.ncu-rep [b.ncu-rep ...] > profiles/ncu_rNN_summary.md"""
import csv, subprocess, sys

METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("tensor pipe active % (DMMA)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("fp64 pipe active %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram read", "dram__bytes_read.sum"), ("dram write", "dram__bytes_write.sum"),
    ("dram throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit rate %", "lts__t_sector_hit_rate.pct"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("stall: barrier", "smsp__pcsamp_warps_issue_stalled_barrier"),
    ("stall: long scoreboard", "smsp__pcsamp_warps_issue_stalled_long_scoreboard"),
    ("stall: short scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard"),
    ("stall: wait (fixed latency)", "smsp__pcsamp_warps_issue_stalled_wait"),
    ("stall: math pipe throttle", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle"),
    ("stall: membar", "smsp__pcsamp_warps_issue_stalled_membar"),
]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print(f"## {path}\n\n(no data)\n")
        continue
    h, units = rows[0], rows[1]
    for rec in rows[2:]:
        name = rec[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        print(f"## {name.split('(')[0].split('::')[-1]}  ({path.split('/')[-1]})\n")
        print("| metric | value |\n|---|---|")
        for label, key in METRICS:
            if key in h:
                i = h.index(key)
                print(f"| {label} (`{key}`) | {rec[i]} {units[i]} |")
        print()
