"""the reference's own L2 callers, UNMODIFIED, on top of this package.

`sbgm/training_utils.py` (get_model), `sbgm/training.py` (TrainingPipeline_general) and
`sbgm/evaluate_sbgm/generation.py` are imported from /root/reference with `sbgm.score_unet` /
`sbgm.score_sampling` redirected by `install_as_sbgm()`.  Their plotting / file-format dependencies that this
image lacks (zarr, netCDF4, matplotlib, omegaconf) are replaced by inert stand-ins; nothing on the model path uses them.

CPU-only container: everything up to the first kernel launch is exercised (module construction from a config
dict, Xavier initialisation through `model.apply`, the `.pth.tar` checkpoint round trip with weights produced by the
REAL reference classes, `train_batches` walking its batch dict into `loss_fn`), and the launch itself must stop at
the no-CPU-fallback RuntimeError.  The same loop with the kernels running is
`test_gpu_train.py::test_reference_epoch_flow_train_validate_generate`.
"""
import importlib.util
import os
import sys
from unittest import mock

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "sbgm")), reason="reference tree not present")

_STUBS = ["zarr", "netCDF4", "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.colors",
          "matplotlib.patches", "matplotlib.dates", "matplotlib.ticker", "matplotlib.cm", "mpl_toolkits",
          "mpl_toolkits.axes_grid1", "omegaconf", "optuna", "cartopy", "seaborn", "cmocean"]


def _cfg(tmp, **model):
    return {
        "experiment": {"config_name": "t"},
        "paths": {"path_save": str(tmp), "checkpoint_dir": str(tmp / "ckpt"), "stats_load_dir": str(tmp / "stats")},
        "highres": {"variable": "temp", "model": "DANRA", "scaling_method": "zscore", "data_size": [32, 32],
                    "full_domain_dims": [589, 789], "cutout_domains": [170, 350, 340, 520]},
        "lowres": {"condition_variables": ["temp", "prcp"], "model": "ERA5", "scaling_methods": ["zscore", "log_zscore"],
                   "data_size": [32, 32], "resize_factor": 1, "full_domain_dims": [589, 789],
                   "cutout_domains": [170, 350, 340, 520]},
        "stationary_conditions": {"geographic_conditions": {"sample_w_geo": True, "geo_variables": ["lsm", "topo"]},
                                  "seasonal_conditions": {"sample_w_cond_season": True, "n_seasons": 5}},
        "sampler": {"time_embedding": 256, "block_layers": [2, 2, 2, 2], "num_heads": 4, "last_fmap_channels": 512,
                    "n_timesteps": 10},
        "transforms": {"scaling": True},
        "training": {"loss_type": "sdfweighted", "weight_init": True, "custom_weight_initializer": None,
                     "sdf_weighted_loss": True, "with_ema": False, "debug_pre_sigma_div": False},
        "model": dict(model),
    }


@pytest.fixture(scope="module")
def ref_l2():
    """(training_utils, training, generation) modules of the reference bound to this package's L1."""
    import sbgm_danra_b200
    saved = {k: v for k, v in sys.modules.items() if k == "sbgm" or k.startswith("sbgm.") or k in _STUBS}
    for k in list(saved):
        sys.modules.pop(k)
    for name in _STUBS:
        m = mock.MagicMock(name=name)
        m.__path__, m.__spec__ = [], None
        sys.modules[name] = m
    sys.path.insert(0, REF)
    try:
        import sbgm  # the reference package itself: its other submodules stay the reference's files  # noqa: F401
        sbgm_danra_b200.install_as_sbgm()
        import sbgm.evaluate_sbgm.generation as gen
        import sbgm.training as tr
        import sbgm.training_utils as tu
        yield tu, tr, gen
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == "sbgm" or k.startswith("sbgm.") or k in _STUBS]:
            sys.modules.pop(k)
        sys.modules.update(saved)


def _real_reference_score_unet():
    spec = importlib.util.spec_from_file_location("_ref_score_unet_for_l2_test", os.path.join(REF, "sbgm", "score_unet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_callers_bind_to_this_package(ref_l2):
    tu, tr, gen = ref_l2
    import sbgm_danra_b200.score_sampling as ss
    import sbgm_danra_b200.score_unet as su
    assert tu.ScoreNet is su.ScoreNet and tu.Encoder is su.Encoder and tu.Decoder is su.Decoder
    assert tr.loss_fn is su.loss_fn and tr.marginal_prob_std_fn is su.marginal_prob_std_fn
    assert tr.Euler_Maruyama_sampler is ss.Euler_Maruyama_sampler and tr.pc_sampler is ss.pc_sampler
    assert gen.pc_sampler is ss.pc_sampler and gen.diffusion_coeff_fn is su.diffusion_coeff_fn
    assert tu.__file__.startswith(REF) and tr.__file__.startswith(REF) and gen.__file__.startswith(REF)


@pytest.mark.parametrize("model_cfg", [dict(), dict(use_resize_conv=False, decoder_norm="instance", decoder_activation="gelu")])
def test_get_model_builds_reference_layout(ref_l2, tmp_path, model_cfg):
    """training_utils.get_model(cfg) (reference :597-669) constructs OUR modules with the reference's state-dict layout."""
    tu, _, _ = ref_l2
    import sbgm_danra_b200.score_unet as su
    cfg = _cfg(tmp_path, **model_cfg)
    model, ckpt_dir, ckpt_name = tu.get_model(cfg)
    assert isinstance(model, su.ScoreNet) and model.debug_pre_sigma_div is False
    assert ckpt_name.endswith(".pth.tar") and ckpt_dir == os.path.join(str(tmp_path), str(tmp_path / "ckpt"))
    ref = _real_reference_score_unet()
    act = {"silu": torch.nn.SiLU, "gelu": torch.nn.GELU}[model_cfg.get("decoder_activation", "SiLU").lower()]
    enc = ref.Encoder(input_channels=6, time_embedding=256, cond_on_img=True, block_layers=[2, 2, 2, 2], num_classes=5, n_heads=4)
    dec = ref.Decoder(last_fmap_channels=512, output_channels=1, time_embedding=256, n_heads=4,
                      use_resize_conv=model_cfg.get("use_resize_conv", True), norm=model_cfg.get("decoder_norm", "group"),
                      gn_groups=8, activation=act)
    theirs = ref.ScoreNet(ref.marginal_prob_std_fn, enc, dec, device="cpu", debug_pre_sigma_div=False).state_dict()
    mine = model.state_dict()
    assert list(mine.keys()) == list(theirs.keys())
    assert all(mine[k].shape == theirs[k].shape and mine[k].dtype == theirs[k].dtype for k in mine)


def test_training_pipeline_checkpoint_round_trip_and_loop_reaches_the_kernels(ref_l2, tmp_path):
    tu, tr, _ = ref_l2
    import sbgm_danra_b200.score_unet as su
    cfg = _cfg(tmp_path)
    cfg["training"]["with_ema"] = True                       # training.py:114 deep-copies the model
    model, _, _ = tu.get_model(cfg)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    pipe = tr.TrainingPipeline_general(model, su.loss_fn, su.marginal_prob_std_fn, su.diffusion_coeff_fn, opt, "cpu", None, cfg)
    # xavier_init_weights (training.py:188-201) went through model.apply: isinstance(nn.Conv2d) must hit our containers
    assert torch.all(model.decoder.final_layer.conv.bias == 0.01)
    assert model.debug_pre_sigma_div is False
    assert isinstance(pipe.ema_model, su.ScoreNet) and pipe.ema_model._cache is not model._cache
    assert torch.equal(pipe.ema_model.encoder.conv1.weight, model.encoder.conv1.weight)

    # checkpoint ABI: weights of the REAL reference classes, saved in the reference's format, load strictly
    ref = _real_reference_score_unet()
    torch.manual_seed(0)
    enc = ref.Encoder(input_channels=6, time_embedding=256, cond_on_img=True, block_layers=[2, 2, 2, 2], num_classes=5, n_heads=4)
    dec = ref.Decoder(last_fmap_channels=512, output_channels=1, time_embedding=256, n_heads=4, use_resize_conv=True,
                      norm="group", gn_groups=8, activation=torch.nn.SiLU)
    theirs = ref.ScoreNet(ref.marginal_prob_std_fn, enc, dec, device="cpu", debug_pre_sigma_div=False)
    path = tmp_path / "ref.pth.tar"
    torch.save({"network_params": theirs.state_dict(), "optimizer_params": {}}, path)
    pipe.load_checkpoint(str(path), device="cpu")
    for k, v in theirs.state_dict().items():
        assert torch.equal(model.state_dict()[k], v), k
    # ... and back: save_model writes a file the reference classes load strictly
    pipe.save_model(dirname=str(tmp_path / "out"), filename="mine.pth")
    theirs.load_state_dict(torch.load(tmp_path / "out" / "mine.pth", map_location="cpu")["network_params"], strict=True)

    # train_batches (training.py:246-422) unpacks the dataset dict and calls loss_fn: on a CPU-only host the call
    # must stop at the kernel boundary, loudly
    g = torch.Generator().manual_seed(0)
    batch = {"temp_hr": torch.randn(2, 1, 32, 32, generator=g), "classifier": torch.randint(1, 5, (2,), generator=g),
             "prcp_lr": torch.randn(2, 1, 32, 32, generator=g), "temp_lr": torch.randn(2, 1, 32, 32, generator=g),
             "lsm": torch.ones(2, 2, 32, 32), "topo": torch.ones(2, 2, 32, 32), "sdf": torch.rand(2, 1, 32, 32, generator=g)}
    if torch.cuda.is_available():
        pytest.skip("kernel launch itself is covered by the GPU suite")
    with pytest.raises(RuntimeError, match="CUDA"):
        pipe.train_batches([batch], epochs=1, current_epoch=1, verbose=False)


class _Attr(dict):
    """Attribute-style access over the nested config dict (what OmegaConf gives generation.py)."""
    def __getattr__(self, k):
        v = self[k]
        return _Attr(v) if isinstance(v, dict) else v


def test_sample_generator_calls_our_pc_sampler(ref_l2, tmp_path):
    """generation.SampleGenerator._run_sampler (reference :56-83) reaches this package's pc_sampler with its keyword
    arguments; on a CPU-only host the call stops at the kernel boundary."""
    tu, _, gen = ref_l2
    cfg = _cfg(tmp_path)
    cfg["paths"]["sample_dir"] = str(tmp_path / "samples")
    model, _, _ = tu.get_model(cfg)
    sg = gen.SampleGenerator(_Attr(cfg), model, dataloader=None, back_transforms=None, device="cpu")
    assert os.path.isdir(sg.sample_path)
    if torch.cuda.is_available():
        pytest.skip("kernel launch itself is covered by the GPU suite")
    with pytest.raises(RuntimeError, match="CUDA"):
        sg._run_sampler(2, torch.tensor([1, 2]), torch.zeros(2, 2, 32, 32), torch.ones(2, 2, 32, 32), torch.ones(2, 2, 32, 32))
