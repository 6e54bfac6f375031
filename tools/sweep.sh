#!/bin/bash
# usage: tools/sweep.sh "<-D flags A>" "<-D flags B>" ...   (run on the GPU box; rebuilds + benches each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
i=0
for flags in "$@"; do
  i=$((i+1))
  touch 3d-reconstruction-from-point-cloud_b200/csrc/pt_knn.cu
  make -C 3d-reconstruction-from-point-cloud_b200 -j8 EXTRA="$flags" > gpurun_out/sweep_build_$i.log 2>&1 || { echo "build failed: $flags"; tail -5 gpurun_out/sweep_build_$i.log; continue; }
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/sweep_$i.log 2>&1
  echo "[$flags] rc=$? $(tail -1 gpurun_out/sweep_$i.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("ms", round(d["ms_per_step"],4), "Msamples/s", round(d["value"]/1e6,1), "e2e", round(d["e2e"]["value"]/1e6,1))' 2>/dev/null)"
done
# leave the default build in place
touch 3d-reconstruction-from-point-cloud_b200/csrc/pt_knn.cu
make -C 3d-reconstruction-from-point-cloud_b200 -j8 > /dev/null 2>&1
