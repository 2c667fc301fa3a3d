"""The FLOP model behind the reported TFLOP/s: `tools/roofline_table.layers` enumerates the forward's operators from the
network's shapes; its totals must reproduce the closed-form figures of SURVEY.md section 2.2 / 8(d) that `bench.py`
(`FWD_FLOP`) and `tools/bench_train.py` divide by."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("size,cin,gflop", [(128, 2, 5.146), (128, 7, 5.313), (64, 2, 1.279), (256, 2, 21.09), (256, 7, 21.76)])
def test_operator_enumeration_reproduces_survey_flops(size, cin, gflop):
    rt = _load("tools/roofline_table.py", "_roofline_table")
    total = sum(fl for _, _, fl, _, _, _ in rt.layers(size, cin)) / 1e9
    assert abs(total - gflop) / gflop < 2e-3, total


def test_bench_constant_matches_the_enumeration():
    rt = _load("tools/roofline_table.py", "_roofline_table")
    bench = _load("bench.py", "_bench_for_flops")
    total = sum(fl for _, _, fl, _, _, _ in rt.layers(bench.SIZE, bench.N_LR + 1))
    assert abs(total - bench.FWD_FLOP) / bench.FWD_FLOP < 1e-3
