#!/bin/bash
# round-2 GPU check #12 (last): training step with the pre-activations kept on the tape (no PReLU-layer re-runs): whole training
# parity suite + step time
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_training.py -q -s > gpurun_out/r02_run12_training.log 2>&1; echo "training rc=$?"
grep -E "passed|failed|greyscale|loss traj|cosine min|^E  |Error" gpurun_out/r02_run12_training.log | head -30
timeout 120 python scripts/train_step_time.py --steps 4 2>&1 | tee gpurun_out/r02_train_step_time_preact.log | tail -3
