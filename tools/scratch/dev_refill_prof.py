"""Old vs refill kernel, profiling builds, host-staged inputs replayed from the device (mixed gaits): per-phase cycles."""
import ctypes, sys, os, subprocess
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
def main(libname):
    import mpcqp, torch
    mpcqp._LIB_PATH = os.path.join(os.path.dirname(mpcqp._LIB_PATH), libname)
    from scenario import Scenario
    lib = mpcqp.load()
    buf = (ctypes.c_ulonglong * 64)()
    for gaits in (["trot", "pace", "bound", "walk"], ["trot"]):
        B, T = 16384, 34
        sc = Scenario(B, gaits=gaits, seed=20260)
        eng = mpcqp.Engine(batch=B)
        hx = np.empty((T, B, 12, 17)); hf = np.empty((T, B, 20, 13))
        for t in range(T):
            xr, fs = sc.inputs(); hx[t], hf[t] = xr, fs
            eng.run(t, xr, fs); x = eng.solution(); sc.advance(x[:, :12] + xr[:, :, 1])
        dx, df = torch.from_numpy(hx).cuda(), torch.from_numpy(hf).cuda()
        eng.reset_warm_start()
        for t in range(25): eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
        eng.synchronize(); lib.mpcqp_debug_profile(buf)
        stream = torch.cuda.ExternalStream(eng.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for t in range(25, T): eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
        e1.record(stream); eng.synchronize()
        lib.mpcqp_debug_profile(buf)
        v = np.array(buf[:], dtype=np.float64)[16:]
        ns, ni, n = max(v[0], 1), max(v[12], 1), (T - 25) * B
        print("%-22s %-20s %.3f ms/tick | warp sweeps per robot pair %.3f, cycles per warp sweep %.0f | per fetch by lane 0: load %.0f decode %.0f finish %.0f (fetches %d for %d pairs)" % (
            libname, "/".join(gaits), e0.elapsed_time(e1) / (T - 25), ns / (n / 2), (v[1] + v[11]) / ns, v[13] / ni, v[14] / ni, v[16] / ni, ni, n // 2), flush=True)
        eng.close()
if __name__ == "__main__":
    if len(sys.argv) > 1: main(sys.argv[1])
    else:
        for l in ("libmpcqp_prof.so", "libmpcqp_prof_refill.so"):
            subprocess.run([sys.executable, __file__, l])
