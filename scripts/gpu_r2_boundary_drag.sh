#!/bin/bash
# One short GPU call: the generic stage kernels after BoundaryDragTerm was added (every other kernel is SASS-identical
# to the library the full GPU suite last passed on): residual parity file (incl. the new boundary-drag tests) first,
# then the Butcher-form / viscous / mode-split steppers as far as the remaining seconds go.
T=gpurun_out/r2q
timeout 40 python -m pytest tests/test_gpu_residual_parity.py tests/test_gpu_erk_steppers.py -q -x -m gpu -p no:cacheprovider > ${T}_tests.txt 2>&1
echo "rc=$?" >> ${T}_tests.txt
tail -5 ${T}_tests.txt
