#!/bin/bash
# round-2 GPU check #8: fused Phong epilogue; backward incl. full-size weight gradients; training step (all-variable gradients,
# Adam trajectory) + its timing; programmatic dependent launch: bit-identity, parity suite under RN_TUNE=pdl=1, same-box A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -s -k 'phong or demo or golden' > gpurun_out/r02_run8_phong.log 2>&1; echo "phong rc=$?"
grep -E "passed|failed|^E  |Error" gpurun_out/r02_run8_phong.log | head -20
timeout 900 python -m pytest tests/test_gpu_training.py -q -s > gpurun_out/r02_run8_training.log 2>&1; echo "training rc=$?"
grep -E "passed|failed|\[exact\]|\[fast\]|loss traj|update of|^E  |Error" gpurun_out/r02_run8_training.log | head -80
timeout 900 python -m pytest tests/test_gpu_backward.py -q -s > gpurun_out/r02_run8_backward.log 2>&1; echo "backward rc=$?"
grep -E "passed|failed|dL/d|wgrad|^E  |Error" gpurun_out/r02_run8_backward.log | head -40
timeout 300 python scripts/wgrad_time.py 2>&1 | tee gpurun_out/r02_wgrad_time.log | tail -12
timeout 600 python scripts/train_step_time.py 2>&1 | tee gpurun_out/r02_train_step_time.log | tail -6
timeout 300 python scripts/pdl_check.py 2>&1 | tee gpurun_out/r02_pdl_check.log; echo "pdl_check rc=$?"
RN_TUNE=pdl=1 timeout 600 python -m pytest tests/test_gpu_exact.py -q -x > gpurun_out/r02_run8_pdl_suite.log 2>&1; echo "pdl suite rc=$?"
tail -3 gpurun_out/r02_run8_pdl_suite.log
timeout 600 python scripts/ab_step.py "base|exact|RN_TUNE=pdl=0" "pdl|exact|RN_TUNE=pdl=1" "base|fast|RN_TUNE=pdl=0" "pdl|fast|RN_TUNE=pdl=1" 2>&1 | tee gpurun_out/r02_pdl_ab.log
