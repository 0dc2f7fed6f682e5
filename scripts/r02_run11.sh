#!/bin/bash
# round-2 GPU check #11: the reference's shipped training configuration (greyscale, BCE) at B = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -q -s -k "greyscale or two_adam" > gpurun_out/r02_run11_training.log 2>&1; echo "training rc=$?"
grep -E "passed|failed|greyscale|loss traj|^E  |Error" gpurun_out/r02_run11_training.log | head -30
