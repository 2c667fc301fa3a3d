"""CPU-only: the inverse-transform oracle (oracle/transforms_ref.py) against golden vectors produced by the reference's own
classes (tests/golden/transforms_golden.npz, make_transforms_golden.py), including the clamp edges and overflow to inf."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import transforms_ref as tr


@pytest.fixture(scope="module")
def tgold():
    return np.load(os.path.join(GOLDEN_DIR, "transforms_golden.npz"))


@pytest.mark.parametrize("name", list(tr.CASES))
def test_oracle_matches_reference_golden(tgold, name):
    x = tgold["x"]
    assert np.array_equal(x, tr.case_input())
    want = tgold[name].astype(np.float64)
    with np.errstate(over="ignore"):
        got = tr.apply_case(name, x)
    fin = np.isfinite(want)
    assert np.array_equal(np.isinf(want), got > 3.4028234663852886e38)       # float32 overflow positions
    # the reference computes in float32: agreement to a few float32 ulp of the log-space value
    assert np.allclose(got[fin], want[fin], rtol=3e-5, atol=1e-6)


def test_constructor_validation_matches_reference():
    with pytest.raises(ValueError):
        tr.prcp_log_back(np.zeros(2), "nope")
