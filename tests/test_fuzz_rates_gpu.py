"""Randomised parity sweep of the rates / credit family against the oracle on the same Philox streams
(tests/fuzz_rates.py): random Vasicek (+ correlated CIR++) models, books of swaps, bonds and a Bermudan swaption,
thresholded and MPoR-collateralised netting sets, metric mixes incl. PFE and CVA, both schemes, with / without Greeks."""
import pytest

import fuzz_rates

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_rate_books_match_the_oracle(seed):
    lines = []
    bad = fuzz_rates.run_cases(12, seed, log=lines.append)
    assert bad == 0, "\n".join(l for l in lines if "MISMATCH" in l)
