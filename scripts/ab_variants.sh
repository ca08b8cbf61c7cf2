#!/bin/bash
# A/B of library builds on the GPU box: scripts/ab_variants.sh NAME... (each _variants/NAME.so, built from a snapshot of csrc/
# with its own -D switches) is copied over the in-tree library and timed with scripts/pipeline_probe.py.
L=riemannian-interior-point-trust-region-method_b200/csrc/libriptrm_b200.so
cp $L /tmp/keep.so
for v in "$@"; do
  cp _variants/$v.so $L
  echo "== $v"
  python scripts/pipeline_probe.py --depths ${DEPTHS:-1,2} --lane 1 --steps ${STEPS:-12} 2>&1 | tail -2
done
cp /tmp/keep.so $L
