#!/bin/bash
# A/B of library builds on what the one-warp-per-copy kernel serves: the lane (monster batch, one step at a time) and small batches
L=riemannian-interior-point-trust-region-method_b200/csrc/libriptrm_b200.so
cp $L /tmp/keep.so
for v in "$@"; do
  cp _variants/$v.so $L
  echo "== $v"
  python scripts/pipeline_probe.py --first-instance 12288 --depths 1,2 --lane 1 --steps 12 2>&1 | tail -2 | cut -c1-30,70-140
  python scripts/pipeline_probe.py --instances 512 --depths 1,2 --lane 1 --steps 24 2>&1 | tail -2 | cut -c1-30,70-140
done
cp /tmp/keep.so $L
