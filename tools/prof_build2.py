"""Per-phase build times (PT_VERBOSE laps) for n points, with and without an id map; three builds
in a row (pool warm).  usage: prof_build2.py [n] [u1]"""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PT_VERBOSE"] = "1"
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
u1 = float(sys.argv[2]) if len(sys.argv) > 2 else 1000.0
w = pkg.synth.CONFIGS["cfg3"]
pos, attrs = pkg.synth.cloud_device(n, w.seed, u1=u1)
ids = torch.arange(n, dtype=torch.int32, device="cuda")
for use_ids in (False, False, True, True):
    print(f"---- n={n} ids={use_ids}", file=sys.stderr, flush=True)
    t = pkg.DeviceTree(pos, attrs, ids if use_ids else None)
    print(f"     build_ms {t.info().build_ms:.2f}", file=sys.stderr, flush=True)
    t.close()
