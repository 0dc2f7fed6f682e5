#!/bin/bash
# round-2 GPU check #4: backward (input-gradient) tests; ncu full capture of the banded res1 conv (half-N MMAs), both precisions
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_backward.py -q -s -x > gpurun_out/r02_run4_backward.log 2>&1; echo "backward rc=$?"
grep -E "passed|failed|rel err|err |cosine|^E  |Error" gpurun_out/r02_run4_backward.log | head -40
for prec in fast exact; do
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm_kernel -s 5 -c 2 -f -o gpurun_out/r02_banded_res1_$prec python scripts/profile_step.py --batch 24 --precision $prec > /dev/null 2>&1; echo "ncu $prec rc=$?"
done
