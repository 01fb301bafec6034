#!/bin/bash
# Evidence for the stage kernel after the instruction-count pass: one `ncu --set full` capture of three config-5 stage
# launches, the default bench line, and (time permitting) the launch list.
T=gpurun_out/r2n
timeout 110 ncu --set full --clock-control none --import-source on -k regex:swe_stage -s 12 -c 3 -o ${T}_swe python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-second-leg > ${T}_ncu_swe.log 2>&1
timeout 170 python bench.py --steps 200 > ${T}_bench_c5_n1.json 2> ${T}_bench_c5_n1.err
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${T}_launches_c5.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-second-leg --no-e2e > ${T}_launches_c5.log 2>&1
