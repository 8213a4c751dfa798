"""N = 64 under commands up to 1 m/s: which robots go to the interior-point stage, tick after tick?"""
import sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario
B, N = 2048, 64
sc = Scenario(B, n_steps=N, gaits=["trot"], seed=4242, noise_kind="hash")
eng = mpcqp.Engine(batch=B, n_steps=N)
eng.scenario_init(sc)
eng.scenario_run(25)
F, S, I = [], [], []
for t in range(30):
    eng.scenario_run(1)
    info = eng.info(with_y=False)
    F.append(info["iters"] > 0); S.append(info["sweeps"].copy()); I.append(info["iters"].copy())
F, S, I = np.array(F), np.array(S), np.array(I)
print("fallback fraction per tick:", np.round(F.mean(axis=1)[:12], 3))
both = (F[1:] & F[:-1]).sum(); print("P(fb_t | fb_t-1) = %.3f   P(fb_t | not fb_t-1) = %.3f" % (both / F[:-1].sum(), (F[1:] & ~F[:-1]).sum() / (~F[:-1]).sum()))
print("robots that fall back on >= 90%% of ticks: %d, never: %d of %d" % ((F.mean(axis=0) >= 0.9).sum(), (F.sum(axis=0) == 0).sum(), B))
print("sweeps: stage-A-only robots mean %.2f; fallback robots: total sweeps mean %.2f, ipm iterations mean %.1f max %d" % (S[~F].mean(), S[F].mean(), I[F].mean(), I.max()))
v = np.abs(sc.current_v_ref()[:, 0])
print("fallback rate by |vx| command quartile:", [round(float(F[:, (v >= a) & (v < b)].mean()), 3) for a, b in ((0, .25), (.25, .5), (.5, .75), (.75, 1.01))])
