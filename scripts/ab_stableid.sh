#!/bin/bash
# A/B of library builds on the StableIdentification sweep: scripts/ab_stableid.sh NAME... (see ab_variants.sh)
L=riemannian-interior-point-trust-region-method_b200/csrc/libriptrm_b200.so
cp $L /tmp/keep.so
for v in "$@"; do
  [ "$v" != "base" ] && cp _variants/$v.so $L
  echo "== $v"
  python scripts/profile_target.py stableid 2>&1 | tail -1 | cut -c1-200
  python scripts/profile_target.py stableid --pairs 16384 2>&1 | tail -1 | cut -c1-200
  cp /tmp/keep.so $L
done
