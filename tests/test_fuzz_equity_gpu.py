"""Randomised parity sweep of the equity family against the oracle on the same Philox streams (tests/fuzz_equity.py):
random Black-Scholes models, books of European / binary / Asian / barrier options, thresholded and MPoR-collateralised
netting sets, metric mixes incl. PFE, with / without a CIR++ counterparty (CVA), with / without exposure Greeks."""
import pytest

import fuzz_equity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_equity_books_match_the_oracle(seed):
    lines = []
    bad = fuzz_equity.run_cases(12, seed, log=lines.append)
    assert bad == 0, "\n".join(l for l in lines if "MISMATCH" in l)
