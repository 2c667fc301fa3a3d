"""CPU: the oracle's restatement of the reference's host->device boundary (oracle/batch_ref.py) against golden vectors made
from the reference's own `extract_samples` and forward transform classes (tests/golden/batch_golden.npz, make_batch_golden.py);
host-side checks of the product's assembler that need no GPU (key rules, errors, the no-CPU-path contract)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import batch_ref as br

NAMES = ("hr", "classifier", "lr", "lsm_hr", "lsm", "sdf", "topo", "hr_point", "lr_point")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN_DIR, "batch_golden.npz"))


@pytest.mark.parametrize("name", list(br.FWD_CASES))
def test_forward_transform_oracle_matches_reference_golden(gold, name):
    x = br.fwd_case_input(br.FWD_CASES[name][2])
    got = br.apply_fwd_case(name, x)
    want = gold[f"fwd/{name}"]
    assert got.dtype == np.float32 and got.shape == want.shape
    # the same float32 operations in the same order; numpy's and torch's log may differ in the last place
    np.testing.assert_allclose(got, want, rtol=3e-7, atol=2.5e-7)     # 1 ulp at O(1): log_minus1_1 cancels near 0


@pytest.mark.parametrize("two_lr", [True, False])
def test_extract_samples_oracle_matches_reference_golden(gold, two_lr):
    res = br.extract_samples_ref(br.sample_dict(two_lr=two_lr))
    for nm, v in zip(NAMES, res):
        want = gold[f"extract{int(two_lr)}/{nm}"]
        assert v is not None
        assert tuple(v.shape) == want.shape and np.array_equal(v.numpy(), want), nm
        if nm != "classifier":
            assert v.dtype == torch.float32
    assert res[2].shape[1] == (3 if two_lr else 1)          # prcp_lr (2 channels) sorted before temp_lr (1)


def test_assembler_has_no_cpu_path_and_keeps_the_reference_errors():
    from sbgm_danra_b200 import batch, special_transforms as st
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        batch.extract_samples(br.sample_dict(), device="cpu")
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        st.ZScoreTransform(1.0, 2.0)(torch.zeros(4))
    with pytest.raises(ValueError, match="Global mean and standard deviation not provided"):
        st.PrcpLogTransform(scale_type="log_zscore")
    with pytest.raises(ValueError, match="Min and max log values not provided"):
        st.PrcpLogTransform(scale_type="log_01")
    with pytest.raises(ValueError, match="Invalid scale type"):
        st.PrcpLogTransform(scale_type="sqrt")
    # kernel parameters restate the reference's arithmetic: (log, eps, sub, mul, div, post_mul, post_add)
    assert st.Scale(-1, 1, 0.0, 160.0).kernel_params() == (False, 0.0, 0.0, 2.0, 160.0, 1.0, -1.0)
    lg = st.PrcpLogTransform(eps=1e-3, scale_type="log_minus1_1", glob_min_log=-6.9, glob_max_log=5.1, buffer_frac=0.25)
    assert lg.kernel_params()[0] is True and lg.kernel_params()[5:] == (2.0, -1.0)
    assert abs(lg.kernel_params()[2] - (-6.9 - 0.25 * 12.0)) < 1e-12 and abs(lg.kernel_params()[4] - 18.0) < 1e-12
