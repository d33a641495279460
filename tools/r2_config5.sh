#!/bin/bash
# BASELINE config 5 on 8 B200: ONE 16384^2 TM solve split into row slabs (halo exchange + allreduce over NCCL) to 1e-10, and the
# modulated MF-FDFD sideband solve on slabs.  Every step under its own timeout.
mkdir -p gpurun_out
L=gpurun_out/r2_config5.log
nvidia-smi -L | wc -l | tee $L
echo "== 16384^2 driven slab solve, 8 ranks" | tee -a $L
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --mode slab --grid 16384 --steps 1 --warmup 0 --no-e2e 2>&1 | grep -E '^\{|rror' | tail -2 | tee gpurun_out/r02_bench_slab_16384_n8.json | cut -c1-1500 | tee -a $L
echo "== modulated slab solve 8192 x 2048 x 3 sidebands, 8 ranks" | tee -a $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/slab_modulated_run.py 8192 2048 2>&1 | grep -E '^\{|rror' | tail -2 | tee gpurun_out/r02_modulated_slab_n8.json | cut -c1-1200 | tee -a $L
