#!/bin/bash
# usage: scripts/gpu_multi.sh N tag   (run under gpurun --gpus N)
N=$1; tag=$2
export OMP_NUM_THREADS=8
if [ "$N" = "2" ]; then
  python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/${tag}_multi_tests.txt 2>&1; echo rc=$? >> gpurun_out/${tag}_multi_tests.txt
fi
for cfg in 5 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $cfg --steps 100 --warmup 5 > gpurun_out/${tag}_bench_c${cfg}_n${N}.json 2> gpurun_out/${tag}_bench_c${cfg}_n${N}.err
done
