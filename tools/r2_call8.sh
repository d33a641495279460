#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call8.log
echo "== gpu suite" | tee $L
timeout 900 python -m pytest tests -m gpu -x -q --timeout 400 2>&1 | tail -15 | tee -a $L
echo "== bench auto" | tee -a $L
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_auto.json | cut -c1-1500 | tee -a $L
echo "== bench bicgstab" | tee -a $L
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --solver bicgstab 2>&1 | tail -1 | tee gpurun_out/bench_bicg.json | cut -c1-1500 | tee -a $L
