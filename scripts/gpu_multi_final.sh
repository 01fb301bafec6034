#!/bin/bash
# usage: scripts/gpu_multi_final.sh N "cfgs"   (run under gpurun --gpus N): driver-style launch of bench.py, both arms
N=$1; cfgs=$2
for cfg in $cfgs; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $cfg --steps 200 --warmup 5 > gpurun_out/r2y_bench_c${cfg}_n${N}.json 2> gpurun_out/r2y_bench_c${cfg}_n${N}.err
done
if [ "$N" = "8" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/r2y_bench_reference_arm_n${N}.json 2> gpurun_out/r2y_bench_reference_arm_n${N}.err
fi
