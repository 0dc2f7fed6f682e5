#!/bin/bash
# round-2 GPU check #7: fused Phong epilogue; backward incl. full-size weight gradients; wgrad kernel timing + ncu;
# programmatic dependent launch: bit-identity, parity suite under RN_TUNE=pdl=1, same-box A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -s -k 'phong or demo or golden' > gpurun_out/r02_run7_phong.log 2>&1; echo "phong rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/r02_run7_phong.log | head -20
timeout 900 python -m pytest tests/test_gpu_backward.py -q -s > gpurun_out/r02_run7_backward.log 2>&1; echo "backward rc=$?"
grep -E "passed|failed|dL/d|wgrad|^E  |Error" gpurun_out/r02_run7_backward.log | head -40
timeout 300 python scripts/wgrad_time.py 2>&1 | tee gpurun_out/r02_wgrad_time.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad2d_kernel -s 3 -c 1 -f -o gpurun_out/r02_wgrad_trunk python scripts/wgrad_time.py > /dev/null 2>&1; echo "ncu wgrad rc=$?"
timeout 300 python scripts/pdl_check.py 2>&1 | tee gpurun_out/r02_pdl_check.log; echo "pdl_check rc=$?"
RN_TUNE=pdl=1 timeout 600 python -m pytest tests/test_gpu_exact.py tests/test_gpu_kernels.py -q -x > gpurun_out/r02_run7_pdl_suite.log 2>&1; echo "pdl suite rc=$?"
tail -5 gpurun_out/r02_run7_pdl_suite.log
timeout 600 python scripts/ab_step.py "base|exact|RN_TUNE=pdl=0" "pdl|exact|RN_TUNE=pdl=1" "base|fast|RN_TUNE=pdl=0" "pdl|fast|RN_TUNE=pdl=1" 2>&1 | tee gpurun_out/r02_pdl_ab.log
