"""A/B of builds, alternating, on the device-resident replay: python tools/scratch/dev_ab3.py libA.so libB.so"""
import os, subprocess, sys
code = ("import sys; sys.path.insert(0,'/root/repo/tools/scratch'); from dev_overlap import run; "
        "run(16384, gaits=['trot','pace','bound','walk'], variants=(1,), ticks=30); run(16384, variants=(1,), ticks=30); run(4096, variants=(1,))")
for rep in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, MPCQP_LIB="/root/repo/mpc-tsid_b200/" + lib)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print("== %-20s " % lib + " | ".join(l.split("tick")[1].split("M solves")[0].strip() for l in r.stdout.splitlines() if l.startswith("B")), flush=True)
