"""Smoke test used by `__graft_entry__.smoke()`: one small score-UNet forward and a 2-step
Euler-Maruyama run on cuda:0, each checked against the CPU oracle (the oracle is the checker only)."""
from __future__ import annotations

import torch


def build_model(cfg, sd, precision: str, device="cuda:0"):
    """ScoreNet of this package for an `synth.NetConfig`, loaded with state-dict `sd`."""
    import torch.nn as nn
    from .score_unet import Decoder, Encoder, ScoreNet, marginal_prob_std_fn
    act = {"relu": nn.ReLU, "silu": nn.SiLU, "gelu": nn.GELU}[cfg.activation]
    enc = Encoder(cfg.in_channels, cfg.time_embedding, block_layers=list(cfg.block_layers), n_heads=cfg.n_heads,
                  num_classes=cfg.num_classes, device=device)
    dec = Decoder(cfg.last_fmap_channels, cfg.out_channels, cfg.time_embedding, n_heads=cfg.n_heads, device=device,
                  use_resize_conv=cfg.use_resize_conv, norm=cfg.norm, gn_groups=cfg.gn_groups, activation=act)
    net = ScoreNet(marginal_prob_std_fn, enc, dec, device=device, debug_pre_sigma_div=False)
    net.load_state_dict(sd, strict=True)
    net.precision = precision
    return net.eval()


def run() -> None:
    from oracle import samplers_ref, score_ref
    from .synth import config_for, synth_batch, synth_state_dict
    from . import score_sampling
    from .score_unet import diffusion_coeff_fn, marginal_prob_std_fn

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    dev = "cuda:0"
    cfg = config_for(n_lr=1)
    sd = synth_state_dict(cfg)
    b = synth_batch(batch=2, size=64, n_lr=1, shared_cond=True)
    with torch.no_grad():
        ref = score_ref.score_forward(sd, cfg, *b.model_args())
    for precision, tol in (("fp32", 1e-4), ("bf16x3", 1e-3), ("fp16x2", 1e-3), ("bf16", 2e-2)):
        net = build_model(cfg, sd, precision, dev)
        with torch.no_grad():
            out = net(b.x.to(dev), b.t.to(dev), None, b.cond_img.to(dev)).cpu()
        err = float((out - ref).norm() / ref.norm())
        print(f"smoke forward [{precision}] rel-L2 vs oracle = {err:.3e} (tol {tol:g})")
        assert err < tol, f"forward parity failed for {precision}: {err}"

    net = build_model(cfg, sd, "bf16x3", dev)
    score_sampling.manual_seed(7)
    got = score_sampling.Euler_Maruyama_sampler(net, marginal_prob_std_fn, diffusion_coeff_fn, batch_size=2, num_steps=2,
                                                device=dev, img_size=64, cond_img=b.cond_img.to(dev)).cpu()
    want = samplers_ref.euler_maruyama(lambda x, t: score_ref.score_forward(sd, cfg, x, t, None, b.cond_img),
                                       score_ref.marginal_prob_std, score_ref.diffusion_coeff, 2, 2, img_size=64,
                                       noise=samplers_ref.philox_noise(7))
    err = float((got - want).norm() / want.norm())
    print(f"smoke EM 2 steps rel-L2 vs oracle = {err:.3e}")
    assert err < 2e-3, f"sampler parity failed: {err}"
    # one DSM training step (loss + backward kernels) against the oracle's autograd on the same Philox draws
    from oracle import philox_ref
    from .score_unet import loss_fn
    cfg_t = config_for(n_lr=1)
    bt = synth_batch(batch=2, size=32, n_lr=1)
    net_t = build_model(cfg_t, sd, "bf16x3", dev).train()
    score_sampling.manual_seed(11)
    loss = loss_fn(net_t, bt.x.to(dev), marginal_prob_std_fn, cond_img=bt.cond_img.to(dev), sdf_cond=bt.sdf_cond.to(dev))
    loss.backward()
    sdo = {k: (v.clone().requires_grad_() if k == "decoder.final_layer.conv.weight" else v.clone()) for k, v in sd.items()}
    u = torch.from_numpy(philox_ref.uniform(2, 11, philox_ref.DRAW_DSM_T))
    z = torch.from_numpy(philox_ref.normal(bt.x.numel(), 11, philox_ref.DRAW_DSM_Z)).reshape(bt.x.shape)
    lo = score_ref.dsm_loss(sdo, cfg_t, bt.x, u * (1.0 - 1e-3) + 1e-3, z, None, bt.cond_img, None, None, bt.sdf_cond, bn_train=True)
    lo.backward()
    g, go = net_t.decoder.final_layer.conv.weight.grad.cpu(), sdo["decoder.final_layer.conv.weight"].grad
    e_loss = abs(loss.item() - lo.item()) / abs(lo.item())
    e_grad = float((g - go).norm() / go.norm())
    print(f"smoke DSM step: loss {loss.item():.5f} vs oracle {lo.item():.5f} (rel {e_loss:.2e}); d(final conv) rel-L2 = {e_grad:.2e}")
    assert e_loss < 1e-3 and e_grad < 2e-3, "training-step parity failed"
    print("smoke OK")
