#!/bin/bash
# Round 2, GPU call 1: run everything that was written blind at the end of round 1, each step under its own timeout.
mkdir -p gpurun_out
L=gpurun_out/r2_call1.log
echo "== 2. multilevel Krylov (unverified)" | tee $L
FDFD_RUN_UNVERIFIED=1 timeout 400 python -m pytest tests/unverified/test_mlkrylov.py -q --timeout 200 2>&1 | tail -40 | tee -a $L
echo "== 3. multilevel Krylov vs default solver on the bench map" | tee -a $L
timeout 500 python tools/gpu_mlkrylov.py 1024 2048 4096 --spec 6,12 6,8 4,8 2>&1 | tee -a $L
echo "== 4c. runtests.jl at exact sizes" | tee -a $L
FDFD_RUN_UNVERIFIED=1 timeout 400 python -m pytest tests/unverified/test_runtests_exact.py -q -s --timeout 300 2>&1 | tail -25 | tee -a $L
echo "== 4b. dolinearsolve seam" | tee -a $L
FDFD_RUN_UNVERIFIED=1 timeout 250 python -m pytest tests/unverified/test_dolinearsolve.py -q --timeout 200 2>&1 | tail -40 | tee -a $L
timeout 250 python tools/gpu_linsolve.py 512 1024 2048 2>&1 | tail -12 | tee -a $L
echo "== 4. slab-sharded modulated / eigenfrequency (unverified)" | tee -a $L
FDFD_RUN_UNVERIFIED=1 timeout 400 python -m pytest tests/unverified/test_slab_multi.py -q --timeout 300 2>&1 | tail -40 | tee -a $L
