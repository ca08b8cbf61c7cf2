#!/bin/bash
# One --set full capture of one kernel of one profile target: bash scripts/ncu_one.sh <tag> <target> <kernel regex> [skip] [count] [extra args]
set -u
TAG=$1; TARGET=$2; KERN=$3; SKIP=${4:-1}; COUNT=${5:-1}; EXTRA=${6:-}
mkdir -p gpurun_out
python scripts/profile_target.py $TARGET $EXTRA > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$KERN -s $SKIP -c $COUNT -o gpurun_out/${TAG} -f \
    python scripts/profile_target.py $TARGET $EXTRA > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page details > gpurun_out/${TAG}_details.txt 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv 2> /dev/null | gzip -9 > gpurun_out/${TAG}_source.csv.gz
rm -f gpurun_out/${TAG}.ncu-rep
cat gpurun_out/${TAG}_plain.log; ls -la gpurun_out | grep ${TAG}
