"""CPU-only: the C-ABI library loads and exports every symbol include/sbgm_b200.h declares (no compute)."""
import os
import re

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "sbgm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sbgm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("sbgm_conv2d_tc", "sbgm_groupnorm", "sbgm_attention", "sbgm_sampler_predictor", "sbgm_dsm_loss",
                 "sbgm_time_embed_project", "sbgm_stem_conv", "sbgm_final_conv", "sbgm_philox_normal"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    from sbgm_danra_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    lib = _lib.load_library()
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/sbgm_b200.h but not exported"
        assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype in sbgm_danra_b200/_lib.py"
    assert set(_lib.PROTOTYPES) == set(_declared())
    assert lib.sbgm_version() >= 100


def test_kernels_are_blackwell_native():
    """The conv kernel's SASS must contain tcgen05 MMA, TMEM loads and TMA loads (B200_PROFILING.md)."""
    import shutil
    import subprocess
    from sbgm_danra_b200 import _lib
    if shutil.which("cuobjdump") is None or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or library unavailable")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
