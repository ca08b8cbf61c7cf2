#!/bin/bash
# the FIXED 16384-pair sweep (BASELINE config 5 as worded) on N GPUs: bash scripts/strong_runs.sh N  (under gpurun --gpus N)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-roofline-leg --scaling strong > gpurun_out/bench_strong$N.json 2> gpurun_out/bench_strong$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-roofline-leg > gpurun_out/bench_weak$N.json 2> gpurun_out/bench_weak$N.err
