"""Profiling driver of the one-robot-per-lane kernel: a few device-resident closed-loop ticks at one robot per lane."""
import sys
import numpy as np
sys.path.insert(0, "/root/repo/mpc-tsid_b200"); sys.path.insert(0, "/root/repo")
import mpcqp
from scenario import Scenario
B = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
sc = Scenario(B, gaits="trot", seed=4242, noise_kind="hash")
eng = mpcqp.Engine(batch=B, mode=29)
eng.scenario_init(sc)
eng.scenario_run(8)
eng.synchronize()
print("ok", eng.info(with_y=False)["sweeps"].mean())
