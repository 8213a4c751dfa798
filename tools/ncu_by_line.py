"""Attribute ncu warp-stall samples to CUDA source lines.
usage: ncu_by_line.py <report.ncu-rep> <libmpcqp.so> <kernel-substring> [launch-index]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line info of the same kernel."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, so, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
# the library links several objects whose embedded cubins share a name: extract each object of the build directory on its own
objs = [so] + sorted(os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", "build", f)
                     for f in os.listdir(os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", "build")) if f.endswith(".o"))
dis = ""
for o in objs:
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(o)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for cub in sorted(os.listdir(tmp)):
        if cub.endswith(".cubin"):
            dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout + "\n"
# per function: list of (line_file, line_no) per instruction in order
funcs, cur, loc = {}, None, ("?", 0)
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        cur = m.group(1); funcs[cur] = []; loc = ("?", 0); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        funcs[cur].append(loc)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sections, cur = [], None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; sections.append(cur); continue
    if cur is not None:
        cur["rows"].append(r)
secs = [s for s in sections if ksub in s["name"]]
sec = secs[which]
H, body = sec["rows"][0], sec["rows"][1:]
if "solve_kernel" in ksub:
    mangled = [f for f in funcs if ("Lb0" in f) == ("(bool)0" in sec["name"]) and "solve_kernel" in f]
elif "riccati_kernel" in ksub:
    n = re.search(r"\(int\)(\d+)", sec["name"]).group(1)
    mangled = [f for f in funcs if "riccati_kernelILi%sELb1" % n in f] or [f for f in funcs if "riccati_kernelILi%sE" % n in f]
else:
    mangled = list(funcs)
lines = funcs[mangled[0]]
si = H.index("# Samples"); ii = H.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
assert len(lines) >= len(body), (len(lines), len(body))
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
tot = 0.0
for r, loc in zip(body, lines):
    s = float(r[si] or 0); tot += s
    a = agg[loc]; a[0] += s; a[1] += float(r[ii] or 0)
    for c in stall_cols:
        v = float(r[c] or 0)
        if v: a[2][H[c]] += v
src_cache = {}
def src(loc):
    f, n = loc
    for d in ("mpc-tsid_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[n - 1].strip()[:90] if 0 < n <= len(L) else ""
    return ""
print("kernel:", sec["name"][:80], " total samples", tot)
allst = collections.Counter()
for a in agg.values(): allst.update(a[2])
print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / sum(allst.values())) for k, v in allst.most_common(8)))
TOP = int(os.environ.get("NCU_BY_LINE_TOP", "40"))
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:TOP]:
    top = ",".join("%s:%.0f%%" % (k[6:], 100 * v / max(1, sum(a[2].values()))) for k, v in a[2].most_common(2))
    print("%5.1f%%  inst %9.0f  %s:%d  [%s]  %s" % (100 * a[0] / tot, a[1], loc[0], loc[1], top, src(loc)))
