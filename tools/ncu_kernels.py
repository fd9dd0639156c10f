"""One markdown table (kernel = column) of the counters DESIGN.md quotes, from an `ncu --set full` report:
    python tools/ncu_kernels.py gpurun_out/r2b_prof.ncu-rep"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
rows = [r for r in rows if r]
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "time us"), ("smsp__inst_executed.sum", "warp inst M"), ("launch__registers_per_thread", "registers"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes per inst"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "pipe fp64 %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe xu %"),
        ("dram__bytes_read.sum", "dram read MB"), ("dram__bytes_write.sum", "dram write MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
idx = {h: i for i, h in enumerate(hdr)}
kn = idx["Kernel Name"]
names = [r[kn].replace("kidmp::", "").split("(")[0].replace("void ", "").replace(", 256", "")[:12] for r in data]


def conv(m, v, u):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if m == "gpu__time_duration.sum":
        x = x / 1e3 if u in ("ns", "nsecond") else (x * 1e3 if u.startswith("ms") else x)
    if m.startswith("dram__bytes"):
        x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
    if m == "smsp__inst_executed.sum":
        x /= 1e6
    return "%.1f" % x


print("| metric | " + " | ".join(names) + " |")
print("|---|" + "---|" * len(names))
for m, label in want:
    if m in idx:
        i = idx[m]
        print("| " + label + " | " + " | ".join(conv(m, r[i], units[i]) for r in data) + " |")
