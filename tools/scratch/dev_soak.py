"""Soak: closed loop on the device, every tick checked: status, pyramid feasibility, sweeps / fallback statistics."""
import sys, time
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario
def soak(N, B, ticks, gaits, seed=7, per_call=1):
    sc = Scenario(B, n_steps=N, gaits=gaits, seed=seed, noise_kind="hash")
    eng = mpcqp.Engine(batch=B, n_steps=N)
    eng.scenario_init(sc)
    bad = 0; msw = 0; mit = 0; fb = 0; worst = 0.0
    t0 = time.time()
    x = np.empty((B, 24 * N))
    for t in range(0, ticks, per_call):
        eng.scenario_run(per_call)                  # per_call >= 2: the ticks of one call overlap as independent index ranges
        st = eng.status()
        bad += int((st != 1).sum())
        if t % 10 == 0 or t >= ticks - per_call:
            info = eng.info(with_y=False); eng.solution(out=x)
            f = x[:, 12 * N:].reshape(B, N, 4, 3)
            mu = eng.params.mu
            v = max((np.abs(f[..., 0]) - mu * f[..., 2]).max(), (np.abs(f[..., 1]) - mu * f[..., 2]).max(), (-f[..., 2]).max(), (f[..., 2] - 25).max())
            worst = max(worst, v); msw = max(msw, info["sweeps"].max()); mit = max(mit, info["iters"].max()); fb += int((info["iters"] > 0).sum())
    print("soak N %d (%d ticks per call): %d robots x %d ticks (%s): not-solved %d, worst pyramid violation %.1e, max sweeps %d, max ipm iters %d, fallback robots on sampled ticks %d, %.1f s"
          % (N, per_call, B, ticks, "/".join(gaits), bad, worst, msw, mit, fb, time.time() - t0), flush=True)
    eng.close()
if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "overlap":
        soak(64, 1024, 100, ["trot"], per_call=2)
        soak(64, 700, 60, ["trot", "pace", "bound", "walk"], per_call=3)
        soak(32, 4096, 100, ["trot", "pace", "bound", "walk"], per_call=2)
        soak(16, 4096, 300, ["trot", "pace", "bound", "walk"], per_call=2)
        soak(16, 6000, 200, ["trot"], per_call=5)
        sys.exit(0)
    soak(64, 1024, 100, ["trot"])
    soak(64, 512, 60, ["trot", "pace", "bound", "walk"])
    soak(32, 1024, 100, ["trot", "pace", "bound", "walk"])
    soak(16, 4096, 300, ["trot", "pace", "bound", "walk"])
    soak(24, 1024, 100, ["trot", "pace", "bound", "walk"])
