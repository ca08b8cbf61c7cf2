#!/bin/bash
# 8-GPU records (run under `gpurun --gpus 8`): weak scaling at 2 and 3 steps in flight, the fixed sweep (strong) at 8 GPUs
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline --no-roofline-leg "${@:3}" > gpurun_out/$2.json 2> gpurun_out/$2.err; }
run 29521 bench_n8_D2
run 29522 bench_n8_D3 --in-flight 3
run 29523 bench_n8_strong --scaling strong
