#!/bin/bash
# The reference's tests/pytests against this package on a GPU box, RNG compatibility mode (its seeded known answers).
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
cd "$ROOT/.reftests_tmp/tests/pytests" || exit 1
PYTHONPATH="$ROOT/tools/reference_suite/stubs:$PYTHONPATH" MCRE_RNG=torch timeout 900 python -m pytest -q . --continue-on-collection-errors 2>&1 | tail -15
