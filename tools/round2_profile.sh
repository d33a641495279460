#!/bin/bash
# Round-2 evidence in the form the measurement contract asks for (B200_PROFILING.md): the ncu launch list of the SAME command as
# the bench line, and one `--set full` capture of the kernel the launch list ranks first among the multigrid kernels.
#   gpurun --timeout 1800 -- 'bash tools/round2_profile.sh'
# Afterwards, here:  python tools/launch_shares.py gpurun_out/r02_bench_launches.csv --md  (share per kernel) and
#   ncu -i gpurun_out/r02_restrict.ncu-rep --page raw --csv | grep -E 'dram__bytes_(read|write)\.sum|gpu__time_duration|l1tex__data_pipe|sm__warps_active'
# then copy the summaries to profiles/r02_*.  Numbers printed by a run under ncu are never bench values.
mkdir -p gpurun_out
echo "== 1. bench line (not under a profiler)" | tee gpurun_out/r2_profile.log
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r02_bench_n1.json | cut -c1-400 | tee -a gpurun_out/r2_profile.log
echo "== 2. launch list of the same command (skip the setup + first iterations, 4000 launches of the steady solve loop)" | tee -a gpurun_out/r2_profile.log
# ncu serialises the four concurrent solves; the step is cut short with --steps 1 --warmup 0 and a hard timeout: only the CSV matters
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30000 -c 4000 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/r02_bench_launches.csv 2>/dev/null)" | tee -a gpurun_out/r2_profile.log
echo "== 3. ncu --set full of k_restrict_tile and k_smooth3 at level 0 (first launches of one solve are level 0)" | tee -a gpurun_out/r2_profile.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_restrict_tile -s 4 -c 1 -o gpurun_out/r02_restrict \
    python tools/prof_solve.py > gpurun_out/r02_restrict.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_smooth3 -s 4 -c 1 -o gpurun_out/r02_smooth3 \
    python tools/prof_solve.py > gpurun_out/r02_smooth3.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tee -a gpurun_out/r2_profile.log
