#!/bin/bash
# fixed 16384-pair sweep on 8 GPUs with more passes in flight: bash scripts/n8_strong_D.sh (under gpurun --gpus 8)
for D in 4 6; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2957$D bench.py --gpus 8 --steps 24 --warmup 6 --in-flight $D --no-cpu-baseline --no-roofline-leg --scaling strong > gpurun_out/bench_n8_strong_D$D.json 2> gpurun_out/bench_n8_strong_D$D.err
done
