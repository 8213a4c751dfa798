"""Device-resident tick time of the one-robot-per-lane kernel against the half-warp kernel at several batch sizes."""
import os, sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import torch, mpcqp
from scenario import Scenario
def run(B, N=16, ticks=12, settle=22, gaits="trot"):
    sc = Scenario(B, n_steps=N, gaits=gaits, seed=4242, noise_kind="hash")
    for name, mode in (("half-warp", 13), ("lane", 29)):
        os.environ["MPCQP_LANE_MIN"] = str(1 << 30)
        eng = mpcqp.Engine(batch=B, n_steps=N, mode=mode)
        eng.scenario_init(sc)
        eng.scenario_run(settle)
        eng.synchronize()
        stream = torch.cuda.ExternalStream(eng.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.scenario_run(ticks)
        eng.join()
        e1.record(stream); eng.synchronize()
        ms = e0.elapsed_time(e1) / ticks
        info = eng.info(with_y=False)
        st = eng.scenario_state()["state"]
        print("B %7d N %d %-9s: %.4f ms/tick  %.2f M solves/s  sweeps %.3f fallback %.4f unsolved %d  checksum %.9f" % (
            B, N, name, ms, B / ms / 1e3, info["sweeps"].mean(), (info["iters"] > 0).mean(), (info["status"] != 1).sum(), np.abs(st).sum()), flush=True)
        eng.close()
if __name__ == "__main__":
    for B in [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else (4096, 65536, 131072):
        run(B)
    run(65536, gaits=["trot", "pace", "bound", "walk"])
    run(16384, N=32, ticks=6)
