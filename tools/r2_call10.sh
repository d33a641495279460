#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call10.log
nvidia-smi -L | tee $L
echo "== NCCL slab test (2 ranks)" | tee -a $L
timeout 900 python -m pytest tests/test_gpu_slab_nccl.py -q --timeout 900 2>&1 | tail -30 | tee -a $L
echo "== bench --gpus 2 (sweep + 8192^2 slab record)" | tee -a $L
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench_n2.json | cut -c1-3000 | tee -a $L
