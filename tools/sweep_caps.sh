cd /root/repo
for flags in "-DPT_TPQ_CAP=8" "-DPT_TPQ_CAP=12"; do
  touch 3d-reconstruction-from-point-cloud_b200/csrc/pt_knn.cu
  make -C 3d-reconstruction-from-point-cloud_b200 -j8 EXTRA="$flags" > /dev/null 2>&1
  echo "== $flags"
  for c in cfg3 cfg4 cfg5 cfg1; do python tools/prof_variants.py $c 2 5 2>&1 | grep variant; done
done
touch 3d-reconstruction-from-point-cloud_b200/csrc/pt_knn.cu; make -C 3d-reconstruction-from-point-cloud_b200 -j8 > /dev/null 2>&1
