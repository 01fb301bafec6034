#!/bin/bash
# Bench lines of the lighter BASELINE configurations with the final stage kernels (no CPU baseline leg: unchanged).
T=gpurun_out/r2n
timeout 40 python bench.py --config 3 --steps 200 --no-cpu-baseline > ${T}_bench_c3_n1.json 2> ${T}_bench_c3_n1.err
timeout 30 python bench.py --config 2 --steps 200 --no-cpu-baseline > ${T}_bench_c2_n1.json 2> ${T}_bench_c2_n1.err
timeout 45 python bench.py --config 4 --steps 200 --no-cpu-baseline > ${T}_bench_c4_n1.json 2> ${T}_bench_c4_n1.err
