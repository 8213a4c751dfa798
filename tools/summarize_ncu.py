"""Summaries for profiles/: (1) per-kernel launch list from `ncu --metrics gpu__time_duration.sum --csv`,
(2) selected metrics of one `ncu --set full` capture.
usage: summarize_ncu.py launches <launches.csv> | metrics <report.ncu-rep>"""
import collections, csv, io, subprocess, sys

def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    H = rows[0]
    kn, mv = H.index("Kernel Name"), H.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[kn], []).append(float(r[mv].replace(",", "")) / 1e3 if "ns" in r[H.index("Metric Unit")] else float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    print("kernel,launches,mean_us,total_ms,share_of_all")
    for k, v in d.items():
        print("%s,%d,%.1f,%.3f,%.3f" % (k.replace(",", ";")[:70], len(v), sum(v) / len(v), sum(v) / 1e3, sum(v) / tot))
    return d

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]

def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H = rows[0]
    print("Kernel Name: " + " | ".join(r[H.index("Kernel Name")] for r in rows[2:]))
    for w in WANT:
        if w in H:
            i = H.index(w)
            print("%s [%s]: %s" % (w, rows[1][i], " | ".join(r[i] for r in rows[2:])))

if __name__ == "__main__":
    {"launches": launches, "metrics": metrics}[sys.argv[1]](sys.argv[2])
