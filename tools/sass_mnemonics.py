"""Static SASS instruction counts per kernel of the built objects (cuobjdump -sass): python tools/sass_mnemonics.py > profiles/rNN_sass_mnemonics.csv
UBLKCP = cp.async.bulk (bulk copy of a robot's xref / gait table into shared memory), SYNCS = mbarrier ops, PREEXIT =
griddepcontrol.launch_dependents, ACQBULK = griddepcontrol.wait, DMMA = FP64 tensor-core MMA (dense solver and its peak probe only)."""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ["UBLKCP", "SYNCS", "DMMA", "DFMA", "DMUL", "DADD", "MUFU", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "ACQBULK", "PREEXIT", "ATOMG", "RED", "WARPSYNC"]
print("# cuobjdump -sass of mpc-tsid_b200/csrc/build/*.o (sm_100a): static instruction counts per kernel (tools/sass_mnemonics.py).")
print("# UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, PREEXIT = griddepcontrol.launch_dependents, ACQBULK = griddepcontrol.wait, DMMA = FP64 tensor-core MMA")
print("kernel,instructions," + ",".join(COLS))
seen = set()
for obj in sorted(glob.glob(os.path.join(ROOT, "mpc-tsid_b200", "csrc", "build", "*.o"))):
    if os.path.basename(obj).startswith("c_"):
        continue
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    name, cnt = None, None
    def flush():
        if name and name not in seen and cnt:
            seen.add(name)
            print("%s,%d,%s" % (name, sum(cnt.values()), ",".join(str(sum(v for k, v in cnt.items() if k.split(".")[0] == c)) for c in COLS)))
    for ln in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            flush()
            dem = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            head = dem[:dem.index(">(") + 1] if ">(" in dem else dem[:dem.index("(")]
            name = head.replace("mpcqp::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "").replace(",", ";")
            cnt = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cnt is not None:
            cnt[m.group(1)] += 1
    flush()
