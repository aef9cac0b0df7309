#!/bin/bash
# Build container only: copies the reference's OWN test files (unmodified) from /root/reference/tests into the
# git-ignored scratch directory .reftests_tmp/ and points their context.py at this package instead of the reference's
# src/.  The copy travels to the GPU box with `gpurun` (it is not gpurun-ignored) but never enters the history.
#   bash tools/reference_suite/make.sh && gpurun -- 'bash tools/reference_suite/run_pytests.sh; bash tools/reference_suite/run_scripts.sh'
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
DST="$ROOT/.reftests_tmp"
rm -rf "$DST/tests"
mkdir -p "$DST"
cp -r /root/reference/tests "$DST/tests"
for d in pytests pv_tests exposure_tests; do
  cat > "$DST/tests/$d/context.py" <<'PY'
import importlib, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", ".."))
sys.path.insert(0, ROOT)
importlib.import_module("montecarlo-risk-engine_b200")
PY
done
echo "reference tests staged under $DST/tests"
