#!/bin/bash
# First GPU call of the next round: confirm the verified suite, then run what was written blind at the end of round 1
# (slab-sharded modulated / eigenfrequency solves, multilevel Krylov solver).  Every step has its own timeout so that a hang
# in unverified code cannot take the box (or the rest of the call) with it.
#   gpurun --timeout 2400 -- 'bash tools/round2_first.sh'
mkdir -p gpurun_out
echo "== 1. verified GPU suite" | tee gpurun_out/r2_first.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee -a gpurun_out/r2_first.log
echo "== 2. multilevel Krylov (unverified)" | tee -a gpurun_out/r2_first.log
FDFD_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests/unverified/test_mlkrylov.py -x -q --timeout 300 2>&1 | tail -15 | tee -a gpurun_out/r2_first.log
echo "== 3. multilevel Krylov vs default solver on the bench map" | tee -a gpurun_out/r2_first.log
timeout 900 python tools/gpu_mlkrylov.py 1024 2048 4096 --spec 6,12 6,8 4,8 2>&1 | tee -a gpurun_out/r2_first.log
echo "== 4. slab-sharded modulated / eigenfrequency (unverified)" | tee -a gpurun_out/r2_first.log
FDFD_RUN_UNVERIFIED=1 timeout 1200 python -m pytest tests/unverified/test_slab_multi.py -x -q --timeout 600 2>&1 | tail -15 | tee -a gpurun_out/r2_first.log
echo "== 4b. dolinearsolve seam on an assembled CSC matrix (unverified)" | tee -a gpurun_out/r2_first.log
FDFD_RUN_UNVERIFIED=1 timeout 600 python -m pytest tests/unverified/test_dolinearsolve.py -x -q --timeout 300 2>&1 | tail -15 | tee -a gpurun_out/r2_first.log
timeout 600 python tools/gpu_linsolve.py 512 1024 2048 2>&1 | tail -12 | tee -a gpurun_out/r2_first.log
echo "== 4c. runtests.jl at exact sizes and isapprox tolerance (unverified sizes)" | tee -a gpurun_out/r2_first.log
FDFD_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests/unverified/test_runtests_exact.py -x -q -s --timeout 400 2>&1 | tail -15 | tee -a gpurun_out/r2_first.log
echo "== 5. bench line with the multilevel solver (only meaningful if step 2 was green)" | tee -a gpurun_out/r2_first.log
timeout 900 python bench.py --solver mlkrylov --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee -a gpurun_out/r2_first.log
