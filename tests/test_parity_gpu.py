"""Parity of the CUDA product with the oracle (tolerances of BASELINE.json's north_star:
max-normalised 1e-10 on kernel blocks, 1e-8 on predictions)."""
import numpy as np
import pytest

from tests._cases import ProductImpl, golden_cases, max_norm_err

pytestmark = pytest.mark.gpu

BLOCK_TOL = 1e-10


@pytest.fixture(scope="module")
def product_outputs(fixtures):
    return golden_cases(fixtures, ProductImpl())


def test_all_golden_cases(product_outputs, golden):
    assert set(product_outputs) == set(golden)
    errs = {k: max_norm_err(product_outputs[k], golden[k]) for k in golden}
    bad = {k: v for k, v in errs.items() if not v <= BLOCK_TOL}
    print("worst max-normalised error: %.3e" % max(errs.values()))
    assert not bad, bad
