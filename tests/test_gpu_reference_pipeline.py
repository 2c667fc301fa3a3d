"""SURVEY section 8(f) rank 1 on the B200: the reference's OWN pipeline objects, unmodified, driving this package's kernels.

`sbgm/training.py::TrainingPipeline_general`, `sbgm/training_utils.py::get_model` and
`sbgm/evaluate_sbgm/generation.py::SampleGenerator` are imported from baseline/_ref (byte-identical copies of the reference's
.py files made by `__graft_entry__.vendor_reference()`; git-ignored, they travel to the GPU box with the snapshot), with
`sbgm.score_unet` / `sbgm.score_sampling` redirected to this package by `install_as_sbgm()` and inert stand-ins for the plotting /
file-format packages the image lacks (zarr, netCDF4, matplotlib, omegaconf; nothing on the model path uses them).

  * `train_batches` (training.py:246-422: extract_samples -> zero_grad -> loss_fn -> backward inside detect_anomaly ->
    optimizer.step -> loss.item()) runs three steps over a synthetic in-memory loader with the dataset's sample-dict schema;
    the three losses are compared with the ORACLE's DSM loss evaluated on the same weights (updated by the same Adam), the
    same batch and the same Philox draws -- not just checked for finiteness;
  * `SampleGenerator._run_sampler` (generation.py:56-83, always pc_sampler) is compared with the oracle's predictor-corrector
    on the same Philox draws; the `.train()` quirk of generation.py:47 is covered by the eval / train pair;
  * the `.pth.tar` checkpoint written by `save_model` on the GPU loads strictly into the reference's own ScoreNet.
"""
import copy
import importlib.util
import os
import sys
from unittest import mock

import pytest
import torch

from conftest import ROOT, rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REF_ROOT = os.path.join(ROOT, "baseline", "_ref")
_STUBS = ["zarr", "netCDF4", "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.colors", "matplotlib.patches",
          "matplotlib.dates", "matplotlib.ticker", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.axes_grid1", "omegaconf", "optuna",
          "cartopy", "seaborn", "cmocean"]


def _cfg(tmp):
    return {
        "experiment": {"config_name": "t"},
        "paths": {"path_save": str(tmp), "checkpoint_dir": str(tmp / "ckpt"), "stats_load_dir": str(tmp / "stats"),
                  "sample_dir": str(tmp / "samples")},
        "highres": {"variable": "temp", "model": "DANRA", "scaling_method": "zscore", "data_size": [32, 32],
                    "full_domain_dims": [589, 789], "cutout_domains": [170, 350, 340, 520]},
        "lowres": {"condition_variables": ["temp", "prcp"], "model": "ERA5", "scaling_methods": ["zscore", "log_zscore"],
                   "data_size": [32, 32], "resize_factor": 1, "full_domain_dims": [589, 789], "cutout_domains": [170, 350, 340, 520]},
        "stationary_conditions": {"geographic_conditions": {"sample_w_geo": True, "geo_variables": ["lsm", "topo"]},
                                  "seasonal_conditions": {"sample_w_cond_season": True, "n_seasons": 4}},
        "sampler": {"time_embedding": 256, "block_layers": [2, 2, 2, 2], "num_heads": 4, "last_fmap_channels": 512, "n_timesteps": 3,
                    "sampler_type": "pc_sampler"},
        "transforms": {"scaling": True},
        "training": {"loss_type": "sdfweighted", "weight_init": False, "custom_weight_initializer": None, "sdf_weighted_loss": True,
                     "with_ema": False, "debug_pre_sigma_div": False, "device": DEV},
        "model": {},
    }


@pytest.fixture(scope="module")
def ref_pipeline():
    if not os.path.isfile(os.path.join(REF_ROOT, "sbgm", "training.py")):
        pytest.skip("baseline/_ref/sbgm absent (run __graft_entry__.build() where /root/reference exists)")
    import sbgm_danra_b200
    saved = {k: v for k, v in sys.modules.items() if k == "sbgm" or k.startswith("sbgm.") or k in _STUBS}
    for k in list(saved):
        sys.modules.pop(k)
    for name in _STUBS:
        m = mock.MagicMock(name=name)
        m.__path__, m.__spec__ = [], None
        sys.modules[name] = m
    sys.path.insert(0, REF_ROOT)
    try:
        import sbgm  # noqa: F401  the reference package; only its two hot-path modules are redirected
        sbgm_danra_b200.install_as_sbgm()
        import sbgm.evaluate_sbgm.generation as gen
        import sbgm.training as tr
        import sbgm.training_utils as tu
        assert tr.__file__.startswith(REF_ROOT) and tu.__file__.startswith(REF_ROOT) and gen.__file__.startswith(REF_ROOT)
        yield tu, tr, gen
    finally:
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k == "sbgm" or k.startswith("sbgm.") or k in _STUBS]:
            sys.modules.pop(k)
        sys.modules.update(saved)


def _oracle_cfg():
    from oracle.synth import config_for
    return config_for(n_lr=2, geo=True, seasons=True)


def _batches(n_batches, batch, size):
    """The sample dicts DANRA_Dataset_cutouts_ERA5_Zarr.__getitem__ yields after collation (data_modules.py:727-997)."""
    g = torch.Generator().manual_seed(11)
    out = []
    for _ in range(n_batches):
        out.append({"temp_hr": torch.randn(batch, 1, size, size, generator=g), "classifier": torch.randint(1, 5, (batch,), generator=g),
                    "temp_lr": torch.randn(batch, 1, size, size, generator=g), "prcp_lr": torch.randn(batch, 1, size, size, generator=g),
                    "lsm": torch.cat([torch.randint(0, 2, (batch, 1, size, size), generator=g).float(), torch.ones(batch, 1, size, size)], 1),
                    "topo": torch.cat([torch.rand(batch, 1, size, size, generator=g) * 2 - 1, torch.ones(batch, 1, size, size)], 1),
                    "sdf": torch.rand(batch, 1, size, size, generator=g)})
    return out


@pytest.mark.parametrize("native_boundary", [False, True])
def test_reference_train_batches_three_steps_match_oracle(ref_pipeline, tmp_path, monkeypatch, native_boundary):
    """native_boundary: the loop's two library calls either side of the kernels are this package's too -- `extract_samples`
    (training.py:287, utils.py:405-480) becomes the one-copy / one-kernel batch assembler and the optimizer the one-launch Adam;
    the reference file still drives every step."""
    from oracle import philox_ref, score_ref
    from sbgm_danra_b200 import batch as sbatch, optim as soptim, score_sampling as ss, score_unet as su
    tu, tr, _ = ref_pipeline
    if native_boundary:
        assert hasattr(tr, "extract_samples")
        monkeypatch.setattr(tr, "extract_samples", sbatch.extract_samples)
    cfg = _cfg(tmp_path)
    model, _, _ = tu.get_model(cfg)                                   # the reference's factory builds THIS package's modules
    assert isinstance(model, su.ScoreNet)
    model = model.to(DEV)
    model.precision = "bf16x3"
    sd0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    opt = (soptim if native_boundary else torch.optim).Adam(model.parameters(), lr=1e-4)
    pipe = tr.TrainingPipeline_general(model, su.loss_fn, su.marginal_prob_std_fn, su.diffusion_coeff_fn, opt, DEV, None, cfg)
    batch, size, steps = 4, 32, 3
    loader = _batches(steps, batch, size)
    losses = []
    orig_loss_fn = tr.loss_fn                                         # train_batches calls the name it imported (training.py:17, :349)
    assert orig_loss_fn is su.loss_fn

    def recording_loss_fn(*a, **kw):
        loss = orig_loss_fn(*a, **kw)
        losses.append(loss.detach().cpu())
        return loss

    monkeypatch.setattr(tr, "loss_fn", recording_loss_fn)             # a recorder around this package's loss_fn; the file is untouched
    ss.manual_seed(700)                                               # step k draws (t, z) from Philox seed 700 + k
    ss.set_ensemble_shard(0, None, None)
    avg = pipe.train_batches(loader, epochs=1, current_epoch=1, verbose=False)
    assert len(losses) == steps and torch.isfinite(torch.stack(losses)).all()
    assert abs(float(avg) - float(torch.stack(losses).mean())) < 1e-3 * abs(float(avg))

    # the oracle, step by step: same weights (torch's Adam on autograd gradients), same batch, same Philox draws
    ocfg = _oracle_cfg()
    params = {k: (v.clone().requires_grad_() if v.is_floating_point() and not k.endswith(("running_mean", "running_var", ".W")) else v.clone())
              for k, v in sd0.items()}
    oopt = torch.optim.Adam([v for v in params.values() if v.requires_grad], lr=1e-4)
    for k, b in enumerate(loader):
        seed = 700 + k
        u = torch.from_numpy(philox_ref.uniform(batch, seed, philox_ref.DRAW_DSM_T))
        z = torch.from_numpy(philox_ref.normal(b["temp_hr"].numel(), seed, philox_ref.DRAW_DSM_Z)).reshape(b["temp_hr"].shape)
        cond = torch.cat([b["prcp_lr"], b["temp_lr"]], 1)            # extract_samples concatenates the *_lr keys in sorted order (utils.py:443)
        oopt.zero_grad()
        lo = score_ref.dsm_loss(params, ocfg, b["temp_hr"], u * (1.0 - 1e-3) + 1e-3, z, b["classifier"], cond, b["lsm"], b["topo"], b["sdf"],
                                bn_train=True)
        lo.backward()
        oopt.step()
        rel = abs(float(losses[k]) - float(lo)) / abs(float(lo))
        print(f"reference train_batches step {k}: loss {float(losses[k]):.5f} vs oracle {float(lo):.5f} (rel {rel:.2e})")
        assert rel < 2e-3, (k, rel)
    # checkpoint written on the GPU by the reference's save_model loads strictly into the reference's own classes
    pipe.save_model(dirname=str(tmp_path / "out"), filename="gpu.pth")
    spec = importlib.util.spec_from_file_location("_ref_su_for_ckpt", os.path.join(REF_ROOT, "sbgm", "score_unet.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    enc = ref.Encoder(input_channels=6, time_embedding=256, cond_on_img=True, block_layers=[2, 2, 2, 2], num_classes=4, n_heads=4)
    dec = ref.Decoder(last_fmap_channels=512, output_channels=1, time_embedding=256, n_heads=4, use_resize_conv=True, norm="group",
                      gn_groups=8, activation=torch.nn.SiLU)
    theirs = ref.ScoreNet(ref.marginal_prob_std_fn, enc, dec, device="cpu", debug_pre_sigma_div=False)
    theirs.load_state_dict(torch.load(tmp_path / "out" / "gpu.pth", map_location="cpu")["network_params"], strict=True)
    assert not torch.equal(theirs.state_dict()["decoder.final_layer.conv.weight"], sd0["decoder.final_layer.conv.weight"])   # it trained


class _Attr(dict):
    def __getattr__(self, k):
        v = self[k]
        return _Attr(v) if isinstance(v, dict) else v


@pytest.mark.parametrize("train_mode", [False, True])
def test_reference_sample_generator_matches_oracle(ref_pipeline, tmp_path, train_mode):
    """SampleGenerator._run_sampler (generation.py:56-83).  Its constructor's `self.model.eval` without parentheses leaves the
    model in whatever mode it was in (:47): both modes are compared with the oracle (train mode = batch-statistics BatchNorm)."""
    from oracle import samplers_ref, score_ref
    from sbgm_danra_b200 import score_sampling as ss
    tu, _, gen = ref_pipeline
    cfg = _cfg(tmp_path)
    model, _, _ = tu.get_model(cfg)
    model = model.to(DEV)
    model.precision = "bf16x3"
    model.train(train_mode)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    sg = gen.SampleGenerator(_Attr(cfg), model, dataloader=None, back_transforms=None, device=DEV)
    assert model.training == train_mode                                # the quirk: no mode change
    b = _batches(1, 4, 32)[0]
    cond = torch.cat([b["temp_lr"], b["prcp_lr"]], 1)
    ss.manual_seed(900)
    ss.set_ensemble_shard(0, None, None)
    got = sg._run_sampler(4, b["classifier"].to(DEV), cond.to(DEV), b["lsm"].to(DEV), b["topo"].to(DEV))
    got = torch.as_tensor(got).reshape(4, 1, 32, 32).float().cpu()
    ocfg = _oracle_cfg()
    score = lambda x, t: score_ref.score_forward(sd, ocfg, x, t, b["classifier"], cond, b["lsm"], b["topo"], bn_train=train_mode)
    with torch.no_grad():
        want = samplers_ref.predictor_corrector(score, score_ref.marginal_prob_std, score_ref.diffusion_coeff, 4, cfg["sampler"]["n_timesteps"],
                                                img_size=32, noise=samplers_ref.philox_noise(900))
    err = rel_l2(got, want)
    print(f"reference SampleGenerator._run_sampler ({'train' if train_mode else 'eval'} mode): rel-L2 vs oracle {err:.2e}")
    assert err < 2e-3
    ss.clear_sampler_cache()
