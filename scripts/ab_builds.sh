for r in 1 2 3; do
  RENDERNET_B200_LIB=gpurun_ab/librn_v2.so python scripts/step_time.py --tag v2_noprefetch_build
  python scripts/step_time.py --tag current
  RN_RES_PREFETCH=0 python scripts/step_time.py --tag current_prefetch_off
  RN_MSUB=1 python scripts/step_time.py --tag current_msub1
done 2>&1 | grep step_time | tee gpurun_out/ab_builds.log
