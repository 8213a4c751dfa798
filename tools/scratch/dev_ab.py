"""A/B of several builds of the library in one box visit: python tools/scratch/dev_ab.py libA.so libB.so ...  (under mpc-tsid_b200/)"""
import os, subprocess, sys
code = "import sys; sys.path.insert(0,'/root/repo/tools/scratch'); from dev_occ import run; run(4096, 0); run(16384, 0)"
for rep in range(2):
    for lib in sys.argv[1:]:
        env = dict(os.environ, MPCQP_LIB="/root/repo/mpc-tsid_b200/" + lib, MPCQP_VERBOSE="1")
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        fit = [l for l in r.stderr.splitlines() if l.startswith("mpcqp:")][:1]
        print("== %-20s " % lib + " | ".join(l.split(":")[1].split("M solves")[0].strip() for l in r.stdout.splitlines() if l.startswith("B")) + "   " + (fit[0][30:] if fit else ""), flush=True)
