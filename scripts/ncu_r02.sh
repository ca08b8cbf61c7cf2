#!/bin/bash
# Round-2 ncu captures (run under gpurun; B200_PROFILING.md recipe).  Outputs land in gpurun_out/.
#   1. launch list of the default bench command (cold-cache, serialised: shares only)
#   2. --set full of the 4 phase launches of sphere_tmem2_kernel in one 16384-pair solve
#   3. --set full of small_kernel<StableIdFam,0> on the 2048-pair sweep
set -u
mkdir -p gpurun_out
TAG=${1:-r02a}
python scripts/profile_target.py sphere > gpurun_out/${TAG}_plain_sphere.log 2>&1 &&
python scripts/profile_target.py stableid > gpurun_out/${TAG}_plain_stableid.log 2>&1 &&
RIPTRM_STABLEID_GENERIC_TCG=1 python scripts/profile_target.py stableid > gpurun_out/${TAG}_plain_stableid_generic.log 2>&1 &&
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain_bench.log 2>&1 || { echo "plain runs failed"; tail -5 gpurun_out/${TAG}_plain_*.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sphere_tmem2 -s 4 -c 4 -o gpurun_out/${TAG}_sphere_tmem2 -f \
    python scripts/profile_target.py sphere > gpurun_out/${TAG}_ncu_sphere.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:small_kernel -s 1 -c 1 -o gpurun_out/${TAG}_stableid -f \
    python scripts/profile_target.py stableid > gpurun_out/${TAG}_ncu_stableid.log 2>&1
RIPTRM_STABLEID_GENERIC_TCG=1 ncu --set full --clock-control none --import-source on -k regex:small_kernel -s 1 -c 1 \
    -o gpurun_out/${TAG}_stableid_generic -f python scripts/profile_target.py stableid > gpurun_out/${TAG}_ncu_stableid_generic.log 2>&1
# keep text exports only: the .ncu-rep files exceed what gpurun copies back (64 MiB)
for name in ${TAG}_sphere_tmem2 ${TAG}_stableid ${TAG}_stableid_generic; do
    if [ -f gpurun_out/$name.ncu-rep ]; then
        ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>&1
        ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>&1
        ncu -i gpurun_out/$name.ncu-rep --page source --csv 2> /dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
        rm -f gpurun_out/$name.ncu-rep
    fi
done
ls -la gpurun_out/ | tail -14
