"""Summarise an .ncu-rep: key raw metrics + hottest SASS lines.  usage: ncu_summary.py rep [n_queries]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; nq = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__sass_inst_executed_op_local_ld.sum','smsp__sass_inst_executed_op_local_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','lts__t_sectors_srcunit_tex_op_read.sum',
        'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size']
for r in rows[2:]:
    print("== kernel:", r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    for i, h in enumerate(hdr):
        if h in want or ('stalled' in h and 'per_issue_active' in h and float(r[i] or 0) > 0.3):
            print(f"  {h} [{units[i]}] = {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ci = h.index("Instructions Executed"); si = h.index("# Samples"); so = h.index("Source")
body = rows[2:]
tot = sum(int(r[ci]) for r in body); tots = sum(int(r[si]) for r in body)
print(f"total warp inst {tot}  per query {tot/nq:.1f}; samples {tots}")
top = sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:40]
for i in sorted(top):
    r = body[i]
    print(f"{i:5d} exec/q {int(r[ci])/nq:8.1f} samples {int(r[si]):6d} ({100*int(r[si])/tots:4.1f}%)  {r[so].strip()[:90]}")
