#!/bin/bash
# full -m gpu suite, then the bench line (not under a profiler); logs under gpurun_out/
mkdir -p gpurun_out
L=gpurun_out/suite_bench.log
echo "== gpu suite" | tee $L
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 2>&1 | tail -8 | tee -a $L
echo "== bench" | tee -a $L
timeout 900 python bench.py --steps ${STEPS:-2} --warmup ${WARMUP:-1} ${BENCH_ARGS} 2>&1 | tail -1 | tee gpurun_out/bench_n1.json | cut -c1-1200 | tee -a $L
