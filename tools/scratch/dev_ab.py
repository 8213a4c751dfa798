"""A/B of several builds of the library in one box visit: python tools/scratch/dev_ab.py libA.so libB.so ...  (under mpc-tsid_b200/)"""
import os, subprocess, sys
code = "import sys; sys.path.insert(0,'/root/repo/tools/scratch'); from dev_occ import run; run(4096, 0); run(16384, 0); run(296, 1)"
for rep in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, MPCQP_LIB="/root/repo/mpc-tsid_b200/" + lib)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout
        print("== %-22s " % lib + " | ".join(l.split(":")[1].split("M solves")[0].strip() for l in out.splitlines() if l.startswith("B")), flush=True)
