#!/bin/bash
# round-2 GPU check #10: thin weight-gradient kernel (e_conv1 / e_conv11): parity, whole training suite, step time after it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -s > gpurun_out/r02_run10_training.log 2>&1; echo "training rc=$?"
grep -E "passed|failed|direct wgrad|loss traj|update of|cosine min|^E  |Error" gpurun_out/r02_run10_training.log | head -60
timeout 600 python scripts/train_step_time.py 2>&1 | tee gpurun_out/r02_train_step_time_thin.log | tail -4
