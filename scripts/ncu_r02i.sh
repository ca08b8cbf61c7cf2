#!/bin/bash
# ncu captures of the round-2 Sphere sweep kernel after the shared-memory reductions / 16-column chunks (run under gpurun).
#   1. launch list of the default bench command (cold-cache, serialised: shares only)
#   2. --set full of the 4 phase launches of sphere_tmem2_kernel in one 16384-pair solve, exported as text
set -u
mkdir -p gpurun_out
TAG=${1:-r02i}
python scripts/profile_target.py sphere > gpurun_out/${TAG}_plain_sphere.log 2>&1 &&
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain_bench.log 2>&1 || { echo "plain runs failed"; tail -5 gpurun_out/${TAG}_plain_*.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sphere_tmem2 -s 4 -c 4 -o gpurun_out/${TAG}_sphere_tmem2 -f \
    python scripts/profile_target.py sphere > gpurun_out/${TAG}_ncu_sphere.log 2>&1
for name in ${TAG}_sphere_tmem2; do
    if [ -f gpurun_out/$name.ncu-rep ]; then
        ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>&1
        ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>&1
        ncu -i gpurun_out/$name.ncu-rep --page source --csv 2> /dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
        rm -f gpurun_out/$name.ncu-rep
    fi
done
ls -la gpurun_out/ | grep ${TAG}
