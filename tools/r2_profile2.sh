#!/bin/bash
# launch list of the bench command (one frequency at a time: ncu serialises kernels anyway; the 4-thread run crashes inside ncu's
# injection library) + level-0 captures of the multigrid kernels picked by grid size + two quick A/B timings
mkdir -p gpurun_out
L=gpurun_out/r2_profile2.log
echo "== launch list" | tee $L
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 12000 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --concurrency 1 --sweep 1 > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/r02_bench_launches.csv 2>/dev/null)" | tee -a $L
tail -2 gpurun_out/r02_bench_under_ncu.log | cut -c1-300 | tee -a $L
echo "== level-0 multigrid kernels (sections, first 40 launches of each family of the live-roofline hook)" | tee -a $L
for k in k_smooth3 k_restrict_tile k_smooth2; do
  timeout 300 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats --section WarpStateStats --section SchedulerStats \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:$k -c 60 --csv --page raw --log-file gpurun_out/r02_mg_$k.csv \
      python -c "
import sys; sys.path.insert(0, '.')
import fdfd_jl_b200 as fdfd
from fdfd_jl_b200 import workloads as wl
d = wl.synthetic_tm_device(fdfd, 4096, 4096, density=1/160)
P = fdfd.Problem(d.grid, fdfd.TM, d.omega[0], d.eps_r, solver=fdfd._lib.SOLVER_BICGSTAB)
print(P.bench_mg({'k_smooth3': 0, 'k_restrict_tile': 1, 'k_smooth2': 2}['$k'], 4))
" > gpurun_out/r02_mg_$k.log 2>&1
  echo "$k rows: $(wc -l < gpurun_out/r02_mg_$k.csv)" | tee -a $L
done
echo "== A/B: rows per CTA of the smoother, heap checker" | tee -a $L
for e in "" "FDFD_MG_S3R=4"; do
  env $e timeout 300 python tools/gpu_r2_exp5.py 4096 2>&1 | head -2 | sed "s/^/[$e] /" | tee -a $L
done
MALLOC_CHECK_=3 timeout 400 python bench.py --steps 1 --warmup 0 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200 | sed "s/^/[MALLOC_CHECK_=3] /" | tee -a $L
