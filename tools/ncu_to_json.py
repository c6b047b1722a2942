"""Summarise an ncu report of decode_stream_kernel into profiles/r02_decode_stream_ncu.json (read by bench.py for roofline.traffic)
and a markdown table:  python tools/ncu_to_json.py gpurun_out/prof_stream_final.ncu-rep <steps_in_launch> <commit>"""
import csv, json, subprocess, sys
rep, steps, commit = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
val = {k: (v[i], u[i]) for i, k in enumerate(h)}
def num(k):
    x, unit = val[k]
    x = float(x.replace(",", ""))
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12, "ms": 1e-3, "us": 1e-6, "s": 1, "ns": 1e-9}.get(unit, 1)
    return x * mult
keys = ["smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "smsp__inst_executed.sum"]
d = {"kernel": "cw::decode_stream_kernel", "report": rep, "commit": commit, "steps_in_launch": steps,
     "duration_s_under_ncu": num("gpu__time_duration.sum"),
     "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum")}
d["dram_bytes_per_launch"] = d["dram_bytes_read"] + d["dram_bytes_write"]
d["dram_bytes_per_step"] = d["dram_bytes_per_launch"] / steps
for k in keys:
    if k in val:
        d[k] = float(val[k][0].replace(",", ""))
json.dump(d, open("profiles/r02_decode_stream_ncu.json", "w"), indent=1)
print(json.dumps(d, indent=1))
