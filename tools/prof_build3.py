"""Device time of index rebuilds (ix->ev events, no verbose laps).  usage: prof_build3.py [n] [reps]"""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w = pkg.synth.CONFIGS["cfg2"]
pos, attrs = pkg.synth.cloud_device(n, w.seed)
for sort in (1, 0):
    pkg.set_option("sort", sort)
    ms = []
    for rep in range(reps):
        t = pkg.DeviceTree(pos, attrs)
        ms.append(t.info().build_ms)
        t.close()
    print(f"sort={sort} build_ms " + " ".join(f"{x:.2f}" for x in ms), flush=True)
