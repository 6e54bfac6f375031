"""Per-phase build times (PT_VERBOSE laps, host wall clock with a stream sync per phase)."""
import os, sys; sys.path.insert(0, "/root/repo")
os.environ["PT_VERBOSE"] = "1"
import torch, __graft_entry__ as ge
pkg = ge.package(); torch.cuda.set_device(0)
w = pkg.synth.CONFIGS["cfg2"]
pos, attrs = pkg.synth.cloud_device(w.n_points, w.seed)
for sort in (1, 0, 1, 0):
    pkg.set_option("sort", sort)
    print(f"---- sort={sort}", file=sys.stderr, flush=True)
    pkg.DeviceTree(pos, attrs).close()
pkg.set_option("sort", 1)
