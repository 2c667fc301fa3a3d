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


def test_oracle_and_package_generators_agree():
    """oracle/synth.py (self-contained: the oracle's network plan does not come from product code) and its product-side twin
    sbgm_danra_b200/synth.py (used by bench.py / tools/, which may not import oracle/) generate the same schema, weights and inputs."""
    import dataclasses

    import torch
    from oracle import synth as o
    from sbgm_danra_b200 import synth as p
    assert o.__file__ != p.__file__ and "sbgm_danra_b200" not in open(o.__file__).read().split('"""', 2)[2]
    for kw in (dict(n_lr=1), dict(n_lr=2, geo=True, seasons=True), dict(n_lr=1, norm="instance", activation="relu", block_layers=(3, 4, 6, 3)),
               dict(n_lr=1, use_resize_conv=False, activation="gelu", n_heads=8)):
        co, cp = o.config_for(**kw), p.config_for(**kw)
        assert dataclasses.asdict(co) == dataclasses.asdict(cp)
        assert list(o.param_schema(co).items()) == list(p.param_schema(cp).items())
        assert o.decoder_plan(co) == p.decoder_plan(cp)
        so, sp = o.synth_state_dict(co, 3), p.synth_state_dict(cp, 3)
        assert list(so) == list(sp) and all(torch.equal(so[k], sp[k]) for k in so)
    bo, bp = o.synth_batch(batch=3, size=32, n_lr=2, geo=True, seasons=True), p.synth_batch(batch=3, size=32, n_lr=2, geo=True, seasons=True)
    for f in ("x", "t", "y", "cond_img", "lsm_cond", "topo_cond", "sdf_cond"):
        assert torch.equal(getattr(bo, f), getattr(bp, f)), f
