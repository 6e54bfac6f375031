#!/bin/bash
# usage: tools/sweep_build.sh "<-D flags A>" ...   rebuilds and reports the index build time (run on the GPU box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for flags in "$@"; do
  touch 3d-reconstruction-from-point-cloud_b200/csrc/*.cu
  make -C 3d-reconstruction-from-point-cloud_b200 -j8 EXTRA="$flags" > gpurun_out/sb.log 2>&1 || { echo "build failed: $flags"; tail -5 gpurun_out/sb.log; continue; }
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/sb_run.log 2>&1
  echo "[$flags] rc=$? $(tail -1 gpurun_out/sb_run.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("build ms", round(d["build"]["ms"],2), "query ms", round(d["ms_per_step"],4))' 2>/dev/null)"
done
touch 3d-reconstruction-from-point-cloud_b200/csrc/*.cu
make -C 3d-reconstruction-from-point-cloud_b200 -j8 > /dev/null 2>&1
