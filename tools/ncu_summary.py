"""Summarise .ncu-rep files (read offline with `ncu -i ... --page raw --csv`) into a markdown table for profiles/."""
import csv, subprocess, sys, io
METRICS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
           ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
           ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
           ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %peak"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
           ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
           ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
           ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
           ("sass__inst_executed_local_loads", "local ld inst")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    cols = [(hdr.index(m), lab) for m, lab in METRICS if m in hdr]
    print(f"\n### {rep.split('/')[-1]}\n")
    print("| kernel | " + " | ".join(lab for _, lab in cols) + " |")
    print("|---|" + "---|" * len(cols))
    ki = hdr.index("Kernel Name")
    for row in rd[2:]:
        vals = []
        for i, lab in cols:
            v = row[i]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {units[i]}".strip())
        print("| `" + row[ki][:48] + "` | " + " | ".join(vals) + " |")
