#!/bin/bash
# First hardware run of TB_OPT_WD_DISPLACED_MASS: the three isolated checks, then the generic-kernel
# parity file as a regression guard.  gpurun --timeout 300 -- 'bash scripts/gpu_wd_displaced_checks.sh'
T=gpurun_out/r3_wd
for c in "stage wetting_drying_alpha_p1" "stage wetting_drying_manning" "refuse" "thacker"; do
  echo "== $c" >> ${T}_checks.txt
  timeout 120 python tests/wd_displaced_gpu_checks.py $c >> ${T}_checks.txt 2>&1
  echo "rc=$?" >> ${T}_checks.txt
done
timeout 120 python -m pytest tests/test_gpu_residual_parity.py -q -x -m gpu -p no:cacheprovider > ${T}_residual_parity.txt 2>&1
echo "rc=$?" >> ${T}_residual_parity.txt
tail -3 ${T}_checks.txt ${T}_residual_parity.txt
