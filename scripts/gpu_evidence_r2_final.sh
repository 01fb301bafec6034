#!/bin/bash
# Round-2 final evidence (one B200): GPU test suite, compute-sanitizer, bench lines of every BASELINE configuration,
# launch lists, ncu captures of the stage / tracer / limiter kernels, dev timings.
set -x
T=gpurun_out/r2z
python -m pytest tests -m gpu -q --durations=8 > ${T}_tests.txt 2>&1; echo "pytest rc=$?" >> ${T}_tests.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > ${T}_sanitizer_memcheck.txt 2>&1; echo "rc=$?" >> ${T}_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > ${T}_sanitizer_racecheck.txt 2>&1; echo "rc=$?" >> ${T}_sanitizer_racecheck.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stepper.py -m gpu -q -x -k "limiter or config4 or two_tracers" > ${T}_sanitizer_memcheck_tracer_limiter.txt 2>&1; echo "rc=$?" >> ${T}_sanitizer_memcheck_tracer_limiter.txt
for c in 1 2 3 4 5; do
  python bench.py --config $c --steps 200 > ${T}_bench_c${c}_n1.json 2> ${T}_bench_c${c}_n1.err
done
python bench.py --impl reference --steps 3 --warmup 1 > ${T}_bench_reference_arm.json 2> ${T}_bench_reference_arm.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${T}_launches_c5.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > ${T}_launches_c5.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${T}_launches_c4.csv python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline > ${T}_launches_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:swe_stage -s 12 -c 3 -o ${T}_swe python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-second-leg > ${T}_ncu_swe.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tracer_stage|limiter" -s 8 -c 4 -o ${T}_tracer python scripts/dev_perf_tracer.py > ${T}_ncu_tracer.log 2>&1
python scripts/dev_perf.py --steps 100 > ${T}_devperf.txt 2>&1
python scripts/dev_perf_tracer.py > ${T}_devperf_tracer.txt 2>&1
python scripts/dev_perf_stommel.py > ${T}_devperf_stommel.txt 2>&1
python scripts/host_overhead.py > ${T}_host_overhead.txt 2>&1
