#!/bin/bash
# The reference's benchmark / plotting scripts (tests/pv_tests, tests/exposure_tests) against this package on a GPU box:
# matplotlib / IPython stubbed, native Philox draws.  One line per script: exit code, seconds, last output line.
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
export PYTHONPATH="$ROOT/tools/reference_suite/stubs:$ROOT/montecarlo-risk-engine_b200:$PYTHONPATH"
cd "$ROOT/.reftests_tmp/tests" || exit 1
for d in pv_tests exposure_tests; do
  for f in $d/*.py; do
    case "$f" in */context.py) continue;; esac
    s=$(date +%s)
    (cd $d && MCRE_RNG=philox timeout 240 python $(basename $f) > /tmp/out.log 2>&1); rc=$?
    e=$(date +%s)
    printf "%s rc=%s %ss :: %s\n" "$f" "$rc" "$((e - s))" "$(grep -v '^\s*$' /tmp/out.log | tail -1 | cut -c1-150)"
  done
done
