#!/bin/bash
# round-2 GPU check #1: exact-mode tests, whole GPU suite, smoke, bench (both precisions)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_run1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_exact.py -q -s -x -k "not config and not stress" > gpurun_out/r02_run1_exact_kernels.log 2>&1; echo "exact kernels rc=$?"
tail -5 gpurun_out/r02_run1_exact_kernels.log
timeout 900 python -m pytest tests/test_gpu_exact.py -q -s -k "config or stress" > gpurun_out/r02_run1_exact_full.log 2>&1; echo "exact full-size rc=$?"
tail -5 gpurun_out/r02_run1_exact_full.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_exact.py > gpurun_out/r02_run1_gpu_suite.log 2>&1; echo "gpu suite rc=$?"
tail -5 gpurun_out/r02_run1_gpu_suite.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_run1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_run1_smoke.log
timeout 900 python bench.py --steps 10 > gpurun_out/r02_run1_bench.json 2> gpurun_out/r02_run1_bench.err; echo "bench rc=$?"
cat gpurun_out/r02_run1_bench.json; tail -5 gpurun_out/r02_run1_bench.err
