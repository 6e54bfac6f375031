#!/bin/bash
# usage: tools/sweep2.sh "<-D flags A>" "<-D flags B>" ...  (GPU box: rebuild + tools/prof_n.py each)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for flags in "$@"; do
  touch 3d-reconstruction-from-point-cloud_b200/csrc/pt_knn.cu
  make -C 3d-reconstruction-from-point-cloud_b200 -j8 EXTRA="$flags" > gpurun_out/sweep2_build.log 2>&1 || { echo "build failed: $flags"; tail -5 gpurun_out/sweep2_build.log; continue; }
  echo "[$flags] $(python ${PROF_SCRIPT:-tools/prof_n.py} $PROF_N_ARGS 2>&1 | tail -${PROF_TAIL:-1})"
done
touch 3d-reconstruction-from-point-cloud_b200/csrc/pt_knn.cu
make -C 3d-reconstruction-from-point-cloud_b200 -j8 > /dev/null 2>&1
