"""Rebuild time vs pool_keep_mb (PT_VERBOSE laps).  usage: prof_pool.py [n]"""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PT_VERBOSE"] = "1"
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
w = pkg.synth.CONFIGS["cfg2"]
pos, attrs = pkg.synth.cloud_device(n, w.seed)
for keep in (2048, 98304, 2048):
    pkg.set_option("pool_keep_mb", keep)
    for rep in range(3):
        print(f"---- keep={keep} rep={rep}", file=sys.stderr, flush=True)
        t = pkg.DeviceTree(pos, attrs)
        print(f"     build_ms {t.info().build_ms:.2f}", file=sys.stderr, flush=True)
        t.close()
