#!/bin/bash
# Every GPU test file under its own timeout (a hung kernel must not hold the box): bash scripts/run_gpu_tests.sh
rc=0
for f in tests/test_*.py; do
  out=$(timeout 900 python -m pytest "$f" -x -q -m gpu 2>&1 | tail -1)
  echo "$f: $out"
  case "$out" in *failed*|*error*|"") rc=1;; esac
done
exit $rc
