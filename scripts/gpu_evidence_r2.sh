#!/bin/bash
# Round-2 evidence run (one B200): GPU test suite, compute-sanitizer, launch list and one full ncu capture.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_tests.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2a_sanitizer_memcheck.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2a_sanitizer_racecheck.txt 2>&1; echo "rc=$?" >> gpurun_out/r2a_sanitizer_racecheck.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:swe_stage -s 12 -c 3 -o gpurun_out/r2a_swe python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1
python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
