set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_scan -s 1 -c 1 -f -o gpurun_out/prof_r1_scan python tools/run_variant.py 5 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:knn_thread -s 1 -c 1 -f -o gpurun_out/prof_r1_final2 python tools/run_variant.py 2 > gpurun_out/ncu_full4.log 2>&1; tail -1 gpurun_out/ncu_full4.log
