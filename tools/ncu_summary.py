"""Summarise an .ncu-rep (raw page) into a small markdown table: python tools/ncu_summary.py rep [rep ...]"""
import csv, io, subprocess, sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma pipe %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
]
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    print(f"\n### {rep}\n")
    print("| kernel | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("<unnamed>::", "")
        cells = []
        for key, _ in COLS:
            if key in hdr:
                i = hdr.index(key)
                cells.append(f"{r[i]} {units[i]}".strip())
            else:
                cells.append("n/a")
        print(f"| `{name}` | " + " | ".join(cells) + " |")
