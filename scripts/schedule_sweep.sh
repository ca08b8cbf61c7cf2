for cfg in "8,14,20:16" "8,20:16" "10,20:16" "20:16" "6,12,19:16" "8,14,21:16" "8,14,20:8" "8,14,20:32" "8,20:32"; do
  sp=${cfg%%:*}; fu=${cfg##*:}
  for f in 0 12288 28672; do
    r=$(RIPTRM_SPLITS=$sp RIPTRM_FAST_UNITS=$fu timeout 120 python scripts/range_probe.py $f | tail -1 | cut -d" " -f3)
    echo -n "$sp/$fu@$f=$r  "
  done; echo
done
