"""Device time of index rebuilds (ix->ev events, no verbose laps) against pool_keep_mb.
usage: prof_build3.py [n] [reps] [keep_mb ...]"""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
keeps = [int(x) for x in sys.argv[3:]] or [-1]
w = pkg.synth.CONFIGS["cfg2"]
pos, attrs = pkg.synth.cloud_device(n, w.seed)
if os.environ.get("PT_SORT"):
    pkg.set_option("sort", int(os.environ["PT_SORT"]))      # 0: cub::DeviceRadixSort (comparison only)
for keep in keeps:
    pkg.set_option("pool_keep_mb", keep)
    ms, wall = [], []
    for rep in range(reps):
        t0 = time.perf_counter()
        t = pkg.DeviceTree(pos, attrs)
        wall.append((time.perf_counter() - t0) * 1e3)
        ms.append(t.info().build_ms)
        t.close()
    print(f"keep={keep} build_ms " + " ".join(f"{x:.2f}" for x in ms) + " | wall " + " ".join(f"{x:.1f}" for x in wall), flush=True)
