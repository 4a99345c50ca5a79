for rep in 1 2; do
  echo "OLD: $(cd _ab/old && python scripts/perf_probe.py --pdl --quick --graph-steps 64 2>&1 | tail -1)"
  echo "NEW: $(python scripts/perf_probe.py --pdl --quick --graph-steps 64 2>&1 | tail -1)"
done
