#!/bin/bash
# Round-2 instruction-count pass on the stage kernels: GPU tests with the new library, then an A/B of the stage kernel
# against the library built from the previous commit (build/ab/lib_0base.so), two passes in opposite order.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -x > gpurun_out/r2m_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/r2m_tests.txt
rm -f gpurun_out/r2m_ab.txt
for lib in build/ab/lib_0base.so build/ab/lib_1new.so build/ab/lib_1new.so build/ab/lib_0base.so; do
  name=$(basename $lib .so)
  THETIS_B200_LIB=$PWD/$lib timeout 200 python scripts/dev_perf.py --mode nonlinear,northsea,northsea_wd --steps 100 2>&1 \
    | grep "ms/stage" | sed "s/^/$name: /" >> gpurun_out/r2m_ab.txt
done
