#!/bin/bash
# Round-2 evidence in the form the measurement contract asks for (B200_PROFILING.md): the bench line, the ncu launch list of the
# SAME bench command, and `--set full` captures of the stencil kernel (DRAM traffic per launch) and of the level-0 multigrid kernels.
#   gpurun --timeout 1800 -- 'bash tools/r2_profile.sh'
# Numbers printed by a run under ncu are never bench values.
mkdir -p gpurun_out
L=gpurun_out/r2_profile.log
echo "== 1. bench line (not under a profiler)" | tee $L
timeout 900 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r02_bench_n1.json | cut -c1-300 | tee -a $L
echo "== 2. launch list of the same command (12000 launches of the steady solve loop)" | tee -a $L
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -s 40000 -c 12000 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/r02_bench_launches.csv 2>/dev/null)" | tee -a $L
echo "== 3. ncu --set full: k_apply (complex128 in, no fused dot: the instantiation bench.py times), k_smooth3, k_restrict_tile at level 0" | tee -a $L
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_apply -s 3 -c 2 -o gpurun_out/r02_k_apply python tools/prof_apply.py 4096 6 > gpurun_out/r02_k_apply.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_restrict_tile -s 6 -c 1 -o gpurun_out/r02_restrict python tools/prof_solve.py 4096 3 > gpurun_out/r02_restrict.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_smooth3 -s 6 -c 1 -o gpurun_out/r02_smooth3 python tools/prof_solve.py 4096 3 > gpurun_out/r02_smooth3.log 2>&1
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tee -a $L
