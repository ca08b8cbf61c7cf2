#!/bin/bash
# 8-GPU records (run under `gpurun --gpus 8`): weak scaling and the fixed sweep (strong), two steps in flight
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-roofline-leg "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; }
run 29521 bench_n8_weak
run 29523 bench_n8_strong --scaling strong
