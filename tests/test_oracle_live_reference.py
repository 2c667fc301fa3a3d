"""The oracle against the LIVE reference (not only the committed goldens): the reference's modules are imported
from /root/reference and evaluated on weights / inputs / option combinations that no golden fixture holds, next to
`oracle.score_ref` on the same tensors.  Build container only; skipped where the reference tree is absent."""
import importlib.util
import os

import pytest
import torch
import torch.nn as nn

from conftest import rel_l2
from oracle import score_ref
from oracle.synth import config_for, synth_batch, synth_state_dict

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "sbgm")), reason="reference tree not present")

ACT = {"relu": nn.ReLU, "silu": nn.SiLU, "gelu": nn.GELU}


@pytest.fixture(scope="module")
def ref_unet():
    spec = importlib.util.spec_from_file_location("_ref_score_unet_live", os.path.join(REF, "sbgm", "score_unet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _reference(ref_unet, cfg, seed):
    enc = ref_unet.Encoder(cfg.in_channels, cfg.time_embedding, block_layers=list(cfg.block_layers), n_heads=cfg.n_heads,
                           num_classes=cfg.num_classes, device="cpu")
    dec = ref_unet.Decoder(cfg.last_fmap_channels, cfg.out_channels, cfg.time_embedding, n_heads=cfg.n_heads, device="cpu",
                           use_resize_conv=cfg.use_resize_conv, norm=cfg.norm, gn_groups=cfg.gn_groups,
                           activation=ACT[cfg.activation])
    net = ref_unet.ScoreNet(ref_unet.marginal_prob_std_fn, enc, dec, device="cpu", debug_pre_sigma_div=False)
    net.load_state_dict(synth_state_dict(cfg, seed), strict=True)
    return net


CASES = {
    "cin7_seasons_gn4_relu": (dict(n_lr=2, geo=True, seasons=True, gn_groups=4, activation="relu"),
                              dict(batch=3, size=32, n_lr=2, geo=True, seasons=True)),
    "cin3_instance_silu_transpose": (dict(n_lr=2, norm="instance", use_resize_conv=False), dict(batch=2, size=32, n_lr=2)),
    "cin5_geo_gelu_heads2_1222": (dict(n_lr=0, geo=True, activation="gelu", n_heads=2, block_layers=(1, 2, 2, 2)),
                                  dict(batch=2, size=64, n_lr=0, geo=True)),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("train", [False, True])
def test_oracle_forward_equals_live_reference(ref_unet, name, train):
    ck, bk = CASES[name]
    cfg = config_for(**ck)
    net = _reference(ref_unet, cfg, seed=5)
    net.train(train)
    b = synth_batch(seed=777, **bk)
    with torch.no_grad():
        want = net(*b.model_args())
        got = score_ref.score_forward(synth_state_dict(cfg, 5), cfg, *b.model_args(), bn_train=train)
    assert rel_l2(got, want) < 2e-5


def test_oracle_dsm_loss_and_gradients_equal_live_reference(ref_unet):
    """loss_fn of the reference (score_unet.py:936-985) with its two random draws pinned, against oracle.dsm_loss: the loss
    and every parameter gradient."""
    ck, bk = CASES["cin7_seasons_gn4_relu"]
    cfg = config_for(**ck)
    net = _reference(ref_unet, cfg, seed=9).train()
    b = synth_batch(seed=31, **bk)
    g = torch.Generator().manual_seed(4)
    u = torch.rand(b.x.shape[0], generator=g)
    z = torch.randn(b.x.shape, generator=g)
    orig = (torch.rand, torch.randn_like)
    torch.rand, torch.randn_like = (lambda *a, **k: u.clone()), (lambda x, **k: z.clone())
    try:
        loss = ref_unet.loss_fn(net, b.x, ref_unet.marginal_prob_std_fn, device="cpu", y=b.y, cond_img=b.cond_img,
                                lsm_cond=b.lsm_cond, topo_cond=b.topo_cond, sdf_cond=b.sdf_cond)
    finally:
        torch.rand, torch.randn_like = orig
    loss.backward()
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in synth_state_dict(cfg, 9).items()}
    lo = score_ref.dsm_loss(sd, cfg, b.x, u * (1.0 - 1e-3) + 1e-3, z, b.y, b.cond_img, b.lsm_cond, b.topo_cond, b.sdf_cond,
                            bn_train=True)
    lo.backward()
    assert abs(lo.item() - loss.item()) / abs(loss.item()) < 1e-5
    checked = 0
    for k, p in net.named_parameters():
        if p.grad is None:
            assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
            continue
        assert rel_l2(sd[k].grad, p.grad) < 1e-3, k
        checked += 1
    assert checked > 150


@pytest.fixture(scope="module")
def ref_samp():
    spec = importlib.util.spec_from_file_location("_ref_score_sampling_live", os.path.join(REF, "sbgm", "score_sampling.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("kind", ["em", "pc"])
def test_oracle_samplers_equal_live_reference_samplers(ref_unet, ref_samp, kind):
    """The UNPATCHED reference samplers at 32x32 (where Euler_Maruyama_sampler's hard-coded initial shape is right,
    score_sampling.py:94) against oracle.samplers_ref with `noise=None`: both draw from torch's global generator at the same
    places, so the same seed gives the same trajectory."""
    from oracle import samplers_ref
    ck, bk = CASES["cin7_seasons_gn4_relu"]
    cfg = config_for(**ck)
    net = _reference(ref_unet, cfg, seed=2).eval()
    sd = synth_state_dict(cfg, 2)
    b = synth_batch(seed=55, **bk)
    kw = dict(batch_size=3, num_steps=5, device="cpu", img_size=32, y=b.y, cond_img=b.cond_img, lsm_cond=b.lsm_cond,
              topo_cond=b.topo_cond)
    score = lambda x, t: score_ref.score_forward(sd, cfg, x, t, b.y, b.cond_img, b.lsm_cond, b.topo_cond)
    torch.manual_seed(123)
    if kind == "em":
        want = ref_samp.Euler_Maruyama_sampler(net, ref_unet.marginal_prob_std_fn, ref_unet.diffusion_coeff_fn, **kw)
    else:
        want = ref_samp.pc_sampler(net, ref_unet.marginal_prob_std_fn, ref_unet.diffusion_coeff_fn, snr=0.16, **kw)
    torch.manual_seed(123)
    fn = samplers_ref.euler_maruyama if kind == "em" else samplers_ref.predictor_corrector
    got = fn(score, score_ref.marginal_prob_std, score_ref.diffusion_coeff, 3, 5, img_size=32)
    assert want.shape == (3, 1, 32, 32) and rel_l2(got, want) < 1e-4


def test_oracle_guided_score_equals_live_reference(ref_unet, ref_samp):
    from oracle import samplers_ref
    ck, bk = CASES["cin7_seasons_gn4_relu"]
    cfg = config_for(**ck)
    net = _reference(ref_unet, cfg, seed=2).eval()
    sd = synth_state_dict(cfg, 2)
    b = synth_batch(seed=56, **bk)
    with torch.no_grad():
        want = ref_samp.guided_score_fn(net, b.x, b.t, b.y, b.cond_img, b.lsm_cond, b.topo_cond, scale=1.5)
        got = samplers_ref.guided_score(lambda *a: score_ref.score_forward(sd, cfg, *a), b.x, b.t, b.y, b.cond_img, b.lsm_cond,
                                        b.topo_cond, scale=1.5)
    assert rel_l2(got, want) < 2e-5


def test_oracle_guided_pc_with_scale_clamp_equals_live_reference(ref_unet, ref_samp):
    """Classifier-free guidance inside pc_sampler with guidance_scale_max < guidance_scale: the reference clamps the scale in
    the corrector only (score_sampling.py:182-186) and re-reads the unclamped one in the predictor (:209-219).  The oracle
    models that with two score callables."""
    from oracle import samplers_ref
    ck, bk = CASES["cin7_seasons_gn4_relu"]
    cfg = config_for(**ck)
    net = _reference(ref_unet, cfg, seed=2).eval()
    sd = synth_state_dict(cfg, 2)
    b = synth_batch(seed=57, **bk)
    guid = {"classifier_free_guidance": {"enabled": True, "guidance_scale": 3.0, "guidance_scale_max": 1.25}}
    torch.manual_seed(321)
    want = ref_samp.pc_sampler(net, ref_unet.marginal_prob_std_fn, ref_unet.diffusion_coeff_fn, snr=0.16, batch_size=3, num_steps=3,
                               device="cpu", img_size=32, y=b.y, cond_img=b.cond_img, lsm_cond=b.lsm_cond, topo_cond=b.topo_cond,
                               cfg=guid)
    model = lambda *a: score_ref.score_forward(sd, cfg, *a)
    guided = lambda w: (lambda x, t: samplers_ref.guided_score(model, x, t, b.y, b.cond_img, b.lsm_cond, b.topo_cond, scale=w))
    torch.manual_seed(321)
    got = samplers_ref.predictor_corrector(guided(1.25), score_ref.marginal_prob_std, score_ref.diffusion_coeff, 3, 3, img_size=32,
                                           score_predictor=guided(3.0))
    assert rel_l2(got, want) < 1e-4
    torch.manual_seed(321)
    wrong = samplers_ref.predictor_corrector(guided(1.25), score_ref.marginal_prob_std, score_ref.diffusion_coeff, 3, 3, img_size=32)
    assert rel_l2(wrong, want) > 1e-3          # one clamped scale for both halves is NOT the reference
