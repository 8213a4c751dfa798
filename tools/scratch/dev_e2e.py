"""End-to-end (host buffers) tick time: kernels reading page-locked host inputs directly against the staged copies (MPCQP_NO_DIRECT=1)."""
import sys, os, time, subprocess
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
def main(B):
    import mpcqp, torch
    from scenario import Scenario
    sc = Scenario(B, gaits="trot", seed=20260)
    eng = mpcqp.Engine(batch=B)
    T = 45
    hx = torch.empty((T, B, 12, 17), dtype=torch.float64, pin_memory=True).numpy()
    hf = torch.empty((T, B, 20, 13), dtype=torch.float64, pin_memory=True).numpy()
    ref = mpcqp.Engine(batch=B)
    for t in range(T):
        xr, fs = sc.inputs(); hx[t], hf[t] = xr, fs
        ref.run(t, xr, fs); x = ref.solution(); sc.advance(x[:, :12] + xr[:, :, 1])       # pageable inputs: staged path
    xlast = x.copy()
    out = torch.empty((B, 12), dtype=torch.float64, pin_memory=True).numpy()
    best = 1e9
    for rep in range(3):
        eng.reset_warm_start()
        for t in range(25):
            eng.run(t, hx[t], hf[t]); eng.forces(out=out)
        t0 = time.perf_counter()
        for t in range(25, T):
            eng.run(t, hx[t], hf[t]); eng.forces(out=out)
        best = min(best, (time.perf_counter() - t0) / (T - 25))
    print("B %d MPCQP_NO_DIRECT=%s: e2e %.3f ms per tick -> %.2f M solves/s, identical to the staged pageable path: %s" % (
        B, os.environ.get("MPCQP_NO_DIRECT", "0"), best * 1e3, B / best * 1e-6, np.array_equal(eng.solution(), xlast)), flush=True)
if __name__ == "__main__":
    if len(sys.argv) > 1:
        main(int(sys.argv[1]))
    else:
        for B in (4096, 16384):
            for nd in ("0", "1"):
                subprocess.run([sys.executable, __file__, str(B)], env=dict(os.environ, MPCQP_NO_DIRECT=nd))
