import sys
sys.path.insert(0, "/root/repo/tools/scratch")
from dev_occ import run
for B in (4096, 16384): run(B, 0)
run(4096, 0, N=32, ticks=40)
run(65536, 0, ticks=40, gaits=["trot", "pace", "bound", "walk"])
