#!/bin/bash
N=$1; tag=$2
export OMP_NUM_THREADS=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config 5 --steps 200 --warmup 5 > gpurun_out/${tag}_bench_c5_n${N}.json 2> gpurun_out/${tag}_bench_c5_n${N}.err
