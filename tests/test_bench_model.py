"""bench.py's roofline inputs, checked on CPU: the useful-FLOP model against a hand count of the algorithm of
mpc-tsid_b200/csrc/mpcqp_riccati.cuh, and the traffic / pipe figures against the committed ncu summary they are parsed from
("traffic matches the ncu summary to the byte")."""
import json
import os

import bench


def test_useful_flops_per_sweep_by_hand():
    # per stage at capacity 16 / 32: two 6x6 factorisations (72 each), T = E U and G = D^-1 + U'T (2 x 126 + 2 x 56),
    # 13 rows x 4 triangular products / substitutions (21 FMA), Pt[:, p] / pt / Pt[:, v] beta for 12 rows, the assembly of P_k
    stage = 2 * 72 + (2 * 126 + 2 * 56) + 13 * 4 * 21 * 2 + (12 * 42 * 2 + 12 * 6 * 2) + (21 * 2 + 36 * 2 + 21 * 5 + 12 * 3)
    for n in (8, 16, 32):
        per_foot = 4 * n * ((18 + 8 + 45 + 66 + 15) + (18 + 18 + 12 + 12 + 20))
        assert bench.flops_stagewise_sweep(n) == n * stage + per_foot + n * (12 * 2 + 6 * 13 * 2 + 6) + n * 36
    # capacity 64 also forms the inverses of the two unit triangular factors: another 72 each
    assert bench.flops_stagewise_sweep(64) - bench.flops_stagewise_sweep(32) * 2 == 64 * 2 * 72
    assert bench.flops_stagewise_sweep(16) == 83984.0


def test_traffic_is_the_committed_ncu_summary_to_the_byte(repo_root):
    prof = "profiles/r02_riccati_kernel_ncu_summary.txt"
    rd = wr = busy = None
    for ln in open(os.path.join(repo_root, prof)):
        if ln.startswith("dram__bytes_read.sum [Mbyte]:"):
            rd = float(ln.split(":")[1]) * 1e6
        if ln.startswith("dram__bytes_write.sum [Mbyte]:"):
            wr = float(ln.split(":")[1]) * 1e6
        if ln.startswith("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"):
            busy = float(ln.split(":")[1])
    assert rd and wr and busy
    assert bench.profile_metric(prof, "dram__bytes_read.sum") == rd
    assert bench.profile_metric(prof, "dram__bytes_write.sum") == wr
    assert bench.profile_metric(prof, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") == busy
    assert bench.profile_metric(prof, "no_such_metric") is None
    # the committed bench line of the round was produced from that file
    line = json.load(open(os.path.join(repo_root, "profiles/r02_bench.json")))
    assert line["roofline"]["traffic"] == rd + wr
    assert abs(line["roofline"]["executed_fp64_pipe_busy_frac"] - busy / 100.0) < 1e-12
    assert line["roofline"]["traffic_source"].find(prof) >= 0
