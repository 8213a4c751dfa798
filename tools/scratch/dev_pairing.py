"""How much of the pair cost max(s1, s2) (two robots share a warp) could pairing by a predictor recover?  Mixed gaits, device closed loop."""
import sys
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "mpc-tsid_b200")); sys.path.insert(0, ROOT)
import mpcqp
from scenario import Scenario
B = 16384
for gaits in (["trot", "pace", "bound", "walk"], ["trot"]):
    sc = Scenario(B, gaits=gaits, seed=4242, noise_kind="hash")
    eng = mpcqp.Engine(batch=B)
    eng.scenario_init(sc)
    eng.scenario_run(25)
    S = []
    for t in range(40):
        eng.scenario_run(1)
        S.append(eng.info(with_y=False)["sweeps"].copy())
    S = np.array(S)                                  # (T, B)
    def pair_cost(s, order):
        o = s[order]
        return np.maximum(o[0::2], o[1::2]).sum() * 2 / len(s)
    kinds = np.array([gaits.index(k) for k in sc.kinds])
    res = {"same gait": [], "same gait + prev tick": [], "same gait + t-16": [], "index": [], "prev tick": [], "prev tick + t-16": [], "t-16": [], "oracle": [], "mean": []}
    for t in range(17, 40):
        s = S[t]
        res["mean"].append(s.mean())
        res["index"].append(pair_cost(s, np.arange(B)))
        res["prev tick"].append(pair_cost(s, np.argsort(S[t - 1], kind="stable")))
        res["t-16"].append(pair_cost(s, np.argsort(S[t - 16], kind="stable")))
        res["prev tick + t-16"].append(pair_cost(s, np.lexsort((S[t - 1], S[t - 16]))))
        res["same gait"].append(pair_cost(s, np.argsort(kinds, kind="stable")))
        res["same gait + prev tick"].append(pair_cost(s, np.lexsort((S[t - 1], kinds))))
        res["same gait + t-16"].append(pair_cost(s, np.lexsort((S[t - 16], kinds))))
        res["oracle"].append(pair_cost(s, np.argsort(s, kind="stable")))
    print("/".join(gaits), {k: round(float(np.mean(v)), 4) for k, v in res.items()})
    c = np.corrcoef(S[17:40].ravel(), S[16:39].ravel())[0, 1]
    c16 = np.corrcoef(S[17:40].ravel(), S[1:24].ravel())[0, 1]
    print("  corr(s_t, s_t-1) %.3f  corr(s_t, s_t-16) %.3f" % (c, c16))
    eng.close()
