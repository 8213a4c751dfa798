"""Memory-safety and race evidence without compute-sanitizer (closed on the B200 pool):

1. the debug build (`make canary` -> libmpcqp_canary.so: guard words between every shared-memory array of a robot and behind
   both halves of its workspace slot, index invariants, all counted on the device) runs every solver path and horizon class
   with zero violations -- and the detector is shown to fire when a kernel is made to write one element too far;
2. racecheck in spirit: results must not depend on how robots are scheduled -- occupancy, batch order, neighbours in the
   warp, batch size -- bit for bit.  A race between the two robots of a warp, between warps sharing a workspace slot, or a read
   of another robot's shared memory would show up as a dependence on exactly these."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANARY = os.path.join(ROOT, "mpc-tsid_b200", "libmpcqp_canary.so")

_SCRIPT = r"""
import ctypes, sys
import numpy as np
sys.path.insert(0, %(root)r + "/mpc-tsid_b200"); sys.path.insert(0, %(root)r)
import mpcqp
from scenario import Scenario
lib = mpcqp.load()
assert hasattr(lib, "mpcqp_debug_canary"), "not the canary build"
def counters(eng):
    out = (ctypes.c_uint * 8)()
    assert lib.mpcqp_debug_canary(eng._h, out) == 0
    return list(out)
total = [0] * 8
for mode, n, B, kw in ((13, 16, 37, {}), (13, 16, 9, {"max_sweeps": 0}), (13, 24, 11, {}), (13, 32, 9, {"max_sweeps": 0}),
                       (13, 64, 5, {}), (13, 64, 4, {"max_sweeps": 1, "ipm_max_iter": 3}), (13, 5, 7, {})):
    dt, T_gait = (0.02, 0.32) if n != 24 else (0.02, 0.48)
    sc = Scenario(B, n_steps=n, dt=dt, T_gait=T_gait, gaits=["trot", "walk", "pace", "bound"], seed=5)
    eng = mpcqp.Engine(batch=B, n_steps=n, mode=mode, dt=dt, T_gait=T_gait, **kw)
    for t in range(4):
        xref, fsteps = sc.inputs()
        eng.run(t, xref, fsteps)
        x = eng.solution()
        assert np.isfinite(x).all()
        sc.advance(x[:, :12] + xref[:, :, 1])
    c = counters(eng)
    total = [a + b for a, b in zip(total, c)]
    eng.close()
sc = Scenario(40, gaits=["trot", "bound", "walk"], seed=9, noise_kind="hash")
eng = mpcqp.Engine(batch=40)
eng.scenario_init(sc)
eng.scenario_run(6)
eng.synchronize()
total = [a + b for a, b in zip(total, counters(eng))]
eng.close()
print("CLEAN", total)
# the detector itself: refine = 77 makes robot 0 write one element past its state array in the debug build
sc = Scenario(4, gaits="trot", seed=1)
eng = mpcqp.Engine(batch=4, refine=77)
xref, fsteps = sc.inputs()
eng.run(0, xref, fsteps)
eng.run(1, xref, fsteps)
eng.synchronize()
print("TRIPPED", counters(eng))
eng.close()
"""


def test_canary_build_is_clean_and_its_detector_fires():
    if not os.path.exists(CANARY):
        pytest.skip("libmpcqp_canary.so not built (make -C mpc-tsid_b200/csrc canary)")
    env = dict(os.environ, MPCQP_LIB=CANARY)
    out = subprocess.run([sys.executable, "-c", _SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = {ln.split()[0]: eval(ln[ln.index("["):]) for ln in out.stdout.splitlines() if ln.startswith(("CLEAN", "TRIPPED"))}
    assert lines["CLEAN"] == [0] * 8, lines
    assert lines["TRIPPED"][0] >= 1 and lines["TRIPPED"][1:] == [0] * 7, lines


@pytest.mark.parametrize("n,kw", [(16, {}), (16, {"max_sweeps": 0}), (24, {}), (64, {})], ids=["N16", "N16-ipm", "N24", "N64"])
def test_results_do_not_depend_on_scheduling(n, kw):
    import mpcqp
    from scenario import Scenario
    B = 301
    dt, T_gait = (0.02, 0.32) if n != 24 else (0.02, 0.48)
    if n == 64:
        B = 93
    sc = Scenario(B, n_steps=n, dt=dt, T_gait=T_gait, gaits=["trot", "pace", "bound", "walk"], seed=41)
    ticks = []
    ref = mpcqp.Engine(batch=B, n_steps=n, dt=dt, T_gait=T_gait, **kw)
    for t in range(3):
        xref, fsteps = sc.inputs()
        ref.run(t, xref, fsteps)
        x = ref.solution()
        ticks.append((xref, fsteps, x.copy(), ref.info()))
        sc.advance(x[:, :12] + xref[:, :, 1])
    ref.close()

    def replay(perm, batch_kw):
        eng = mpcqp.Engine(batch=len(perm), n_steps=n, dt=dt, T_gait=T_gait, **kw, **batch_kw)
        for t, (xref, fsteps, x, info) in enumerate(ticks):
            eng.run(t, np.ascontiguousarray(xref[perm]), np.ascontiguousarray(fsteps[perm]))
            got, gi = eng.solution(), eng.info()
            np.testing.assert_array_equal(got, x[perm])
            np.testing.assert_array_equal(gi["sweeps"], info["sweeps"][perm])
            np.testing.assert_array_equal(gi["iters"], info["iters"][perm])
            np.testing.assert_array_equal(gi["y"], info["y"][perm])
        eng.close()

    rng = np.random.default_rng(0)
    replay(np.arange(B)[::-1].copy(), {})                      # other neighbours in every warp
    replay(rng.permutation(B), {})
    replay(np.arange(0, B, 7), {})                             # another batch size: other robots, other tail
    os.environ["MPCQP_RIC_CTAS"] = "1"                         # one warp per SM: another interleaving of everything
    try:
        replay(np.arange(B), {})
    finally:
        del os.environ["MPCQP_RIC_CTAS"]


@pytest.mark.parametrize("n,ranges,kw", [(16, 2, {}), (16, 3, {}), (16, 8, {}), (16, 4, {"max_sweeps": 0}), (32, 8, {}), (64, 2, {})],
                         ids=["N16-2", "N16-3", "N16-8", "N16-4-ipm", "N32-8", "N64-2"])
def test_overlapped_index_ranges_change_nothing(n, ranges, kw):
    """mpcqp_set_overlap: consecutive device-resident ticks issued as independent index ranges (a range's tick t + 1 is ordered
    behind its own tick t only, each range has its own fallback queue and workspace) must give bit-identical solutions,
    multipliers, sweep and iteration counts, fallback counts and world poses, tick after tick -- for host-staged inputs replayed
    from the device and for the device-resident closed loop."""
    import torch
    import mpcqp
    from scenario import Scenario
    B = 1201 if n == 16 else (333 if n == 32 else 95)
    gaits = ["trot", "pace", "bound", "walk"]
    sc = Scenario(B, n_steps=n, gaits=gaits, seed=43)
    ref = mpcqp.Engine(batch=B, n_steps=n, **kw)
    ticks = []
    for t in range(6):
        xref, fsteps = sc.inputs()
        ref.run(t, xref, fsteps)
        x = ref.solution()
        ticks.append((xref, fsteps, x.copy(), ref.info(), ref.fallback_count(), ref.world_pose()))
        sc.advance(x[:, :12] + xref[:, :, 1])
    ref.close()
    dx = [torch.from_numpy(tk[0]).cuda() for tk in ticks]
    df = [torch.from_numpy(tk[1]).cuda() for tk in ticks]
    torch.cuda.synchronize()
    eng = mpcqp.Engine(batch=B, n_steps=n, **kw)
    eng.set_overlap(ranges)
    # all six ticks in flight at once, then the last one's results; then tick by tick
    for t in range(6):
        eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
    np.testing.assert_array_equal(eng.solution(), ticks[5][2])
    np.testing.assert_array_equal(eng.world_pose(), ticks[5][5])
    assert eng.fallback_count() == ticks[5][4]
    eng.reset_warm_start()
    eng.world_pose(set_to=np.tile([0, 0, 0.2027682, 0, 0, 0], (B, 1)))
    for t, (xref, fsteps, x, info, nfb, qw) in enumerate(ticks):
        eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
        got, gi = eng.solution(), eng.info()
        np.testing.assert_array_equal(got, x)
        for key in ("status", "sweeps", "iters", "y", "obj", "active", "contact"):
            np.testing.assert_array_equal(gi[key], info[key])
        assert eng.fallback_count() == nfb
    # switching back to host inputs joins the ranges first
    eng.set_overlap(ranges)
    eng.reset_warm_start()
    for t in range(3):
        eng.run_device(t, dx[t].data_ptr(), df[t].data_ptr())
    for t in range(3, 6):
        eng.run(t, ticks[t][0], ticks[t][1])
    np.testing.assert_array_equal(eng.solution(), ticks[5][2])
    # host inputs overlap too (every range stages its own rows on its own stream); the asynchronous result protocol copies every
    # range's forces behind its own tick: issue tick t + 1, then wait for the forces of tick t
    eng.set_overlap(ranges)
    eng.reset_warm_start()
    px = [torch.from_numpy(tk[0]).pin_memory().numpy() for tk in ticks]
    pf = [torch.from_numpy(tk[1]).pin_memory().numpy() for tk in ticks]
    for t in range(6):
        eng.run(t, px[t], pf[t])
        eng.result_async(t & 1)
        if t > 0:
            np.testing.assert_array_equal(eng.result_wait((t - 1) & 1), ticks[t - 1][2][:, 12 * n:12 * n + 12])
    np.testing.assert_array_equal(eng.result_wait(5 & 1), ticks[5][2][:, 12 * n:12 * n + 12])
    np.testing.assert_array_equal(eng.solution(), ticks[5][2])
    eng.close()

    # device-resident closed loop: one call of 8 ticks (ranges) against 8 calls of one tick (never overlapped)
    res = []
    for how in ("single", "ranges"):
        scd = Scenario(B, n_steps=n, gaits=gaits, seed=44, noise_kind="hash")
        e = mpcqp.Engine(batch=B, n_steps=n, **kw)
        e.scenario_init(scd)
        if how == "single":
            e.set_overlap(1)
            for _ in range(8):
                e.scenario_run(1)
        else:
            e.set_overlap(ranges)
            e.scenario_run(5)
            e.scenario_run(3)
        st = e.scenario_state()
        res.append((e.solution(), st["state"], st["frame"], st["feet"], e.info()["sweeps"]))
        e.close()
    for a, b in zip(res[0], res[1]):
        np.testing.assert_array_equal(a, b)
