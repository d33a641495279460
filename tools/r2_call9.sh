#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_call9.log
echo "== eigen + parity" | tee $L
timeout 1500 python -m pytest tests/test_gpu_modulated_eigen.py tests/test_gpu_slab_multi.py "tests/test_gpu_parity.py::test_bench_map_vs_oracle_direct_solve" tests/test_gpu_parity.py::test_config2_directional_coupler_vs_direct_solve tests/test_gpu_mlkrylov.py::test_eigenfrequency_with_multilevel_inner_solves -q --timeout 900 --durations=12 2>&1 | tail -40 | tee -a $L
free -g | head -2 | tee -a $L
