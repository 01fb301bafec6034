#!/bin/bash
set -x
python -m pytest tests/test_gpu_stepper.py tests/test_gpu_erk_steppers.py -m gpu -q -x > gpurun_out/r2e_tests.txt 2>&1; echo rc=$? >> gpurun_out/r2e_tests.txt
python scripts/profile_e2e.py 7 graphs fused > gpurun_out/r2e_prof_e2e_k7.txt 2>&1
python bench.py --steps 100 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:swe_stage -s 12 -c 3 -o gpurun_out/r2e_swe python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-second-leg > gpurun_out/r2e_ncu.log 2>&1
