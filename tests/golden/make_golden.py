"""Generate the golden fixtures in this directory from the REAL reference.

Run in the build container only (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules are imported unmodified from /root/reference.  Weights come from
`oracle.synth.synth_state_dict` (loaded with `load_state_dict(strict=True)`, which also pins
the key/shape schema), inputs from `oracle.synth.synth_batch`, and every random draw the
reference makes (`torch.randn`, `torch.randn_like`, `torch.rand`) is replaced for the duration
of the call by the Philox stream of `oracle.philox_ref`, so the fixtures are reproducible
and the CUDA kernels can consume the very same noise.
"""
from __future__ import annotations

import contextlib
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import philox_ref  # noqa: E402
from oracle.synth import NetConfig, config_for, synth_batch, synth_state_dict  # noqa: E402

from sbgm import score_sampling as ref_samp  # noqa: E402
from sbgm import score_unet as ref_unet  # noqa: E402

ACT = {"relu": nn.ReLU, "silu": nn.SiLU, "gelu": nn.GELU}
SEED_NOISE = 2024


def build_reference(cfg: NetConfig, seed: int = 0):
    enc = ref_unet.Encoder(cfg.in_channels, cfg.time_embedding, block_layers=list(cfg.block_layers),
                           n_heads=cfg.n_heads, num_classes=cfg.num_classes, device="cpu")
    dec = ref_unet.Decoder(cfg.last_fmap_channels, cfg.out_channels, cfg.time_embedding, n_heads=cfg.n_heads,
                           device="cpu", use_resize_conv=cfg.use_resize_conv, norm=cfg.norm,
                           gn_groups=cfg.gn_groups, activation=ACT[cfg.activation])
    net = ref_unet.ScoreNet(ref_unet.marginal_prob_std_fn, enc, dec, device="cpu", debug_pre_sigma_div=False)
    net.load_state_dict(synth_state_dict(cfg, seed), strict=True)
    return net


@contextlib.contextmanager
def injected_noise(queue):
    """Replace torch's global normal/uniform draws by the next tensors of `queue` (callables shape->tensor)."""
    orig = (torch.randn, torch.randn_like, torch.rand)
    it = iter(queue)

    def randn(*a, **k):
        return next(it)("randn")

    def randn_like(x, **k):
        return next(it)("randn_like").reshape(x.shape)

    def rand(*a, **k):
        return next(it)("rand")

    torch.randn, torch.randn_like, torch.rand = randn, randn_like, rand
    try:
        yield
    finally:
        torch.randn, torch.randn_like, torch.rand = orig


def normal_draw(draw, shape):
    return lambda kind: torch.from_numpy(philox_ref.normal(int(np.prod(shape)), SEED_NOISE, draw)).reshape(shape)


def uniform_draw(draw, shape):
    return lambda kind: torch.from_numpy(philox_ref.uniform(int(np.prod(shape)), SEED_NOISE, draw)).reshape(shape)


FORWARD_CASES = {
    # name: (config kwargs, batch kwargs)
    "fwd_c1_64_cin2": (dict(n_lr=1), dict(batch=2, size=64, n_lr=1)),
    "fwd_c3_64_cin7_seasons": (dict(n_lr=2, geo=True, seasons=True), dict(batch=2, size=64, n_lr=2, geo=True, seasons=True)),
    "fwd_32_instance_relu_3463": (dict(n_lr=1, norm="instance", activation="relu", block_layers=(3, 4, 6, 3)),
                                  dict(batch=2, size=32, n_lr=1)),
    "fwd_32_transpose_gelu_h8": (dict(n_lr=1, use_resize_conv=False, activation="gelu", n_heads=8),
                                 dict(batch=3, size=32, n_lr=1)),
    "fwd_128_cin2": (dict(n_lr=1), dict(batch=1, size=128, n_lr=1)),
}


def main() -> None:
    torch.set_num_threads(8)
    out = {}
    schema = {}
    for name, (ck, bk) in FORWARD_CASES.items():
        cfg = config_for(**ck)
        net = build_reference(cfg).eval()
        schema[name] = {k: list(v.shape) for k, v in net.state_dict().items()}
        b = synth_batch(**bk)
        with torch.no_grad():
            fmaps = net.encoder(b.x, b.t, y=b.y, cond_img=b.cond_img, lsm_cond=b.lsm_cond, topo_cond=b.topo_cond)
            score = net(*b.model_args())
        out[f"{name}/score"] = score.numpy()
        out[f"{name}/fmap_stats"] = np.array([[f.mean().item(), f.std().item()] for f in fmaps], dtype=np.float64)
        print(name, "score rms", score.pow(2).mean().sqrt().item())
        if name == "fwd_c3_64_cin7_seasons":
            net.train()   # BatchNorm with batch statistics (generation.py:47 quirk, training)
            with torch.no_grad():
                out[f"{name}/score_bn_train"] = net(*b.model_args()).numpy()
            net.eval()

    # --- samplers, 32x32 and 64x64, 3 steps, Philox-injected noise -------------------------------
    for name, size, ck, bk in (("c1", 64, dict(n_lr=1), dict(n_lr=1)),
                               ("c3", 32, dict(n_lr=2, geo=True, seasons=True), dict(n_lr=2, geo=True, seasons=True))):
        cfg = config_for(**ck)
        net = build_reference(cfg).eval()
        B, N = 2, 3
        b = synth_batch(batch=B, size=size, shared_cond=True, **bk)
        shape = (B, 1, size, size)
        kw = dict(batch_size=B, num_steps=N, device="cpu", eps=1e-3, img_size=size, y=b.y, cond_img=b.cond_img,
                  lsm_cond=b.lsm_cond, topo_cond=b.topo_cond)
        q = [normal_draw(philox_ref.DRAW_INIT, shape)] + [normal_draw(philox_ref.draw_em(k), shape) for k in range(N)]
        with injected_noise(q):
            em = ref_samp.Euler_Maruyama_sampler(net, ref_unet.marginal_prob_std_fn, ref_unet.diffusion_coeff_fn, **kw)
        out[f"em_{name}/mean_x"] = em.numpy()
        q = [normal_draw(philox_ref.DRAW_INIT, shape)]
        for k in range(N):
            q += [normal_draw(philox_ref.draw_pc_corrector(k), shape), normal_draw(philox_ref.draw_pc_predictor(k), shape)]
        with injected_noise(q):
            pc = ref_samp.pc_sampler(net, ref_unet.marginal_prob_std_fn, ref_unet.diffusion_coeff_fn, snr=0.16, **kw)
        out[f"pc_{name}/x_mean"] = pc.numpy()
        if name == "c3":
            cfgd = {"classifier_free_guidance": {"enabled": True, "guidance_scale": 1.5}}
            q = [normal_draw(philox_ref.DRAW_INIT, shape)] + [normal_draw(philox_ref.draw_em(k), shape) for k in range(N)]
            with injected_noise(q):
                emg = ref_samp.Euler_Maruyama_sampler(net, ref_unet.marginal_prob_std_fn,
                                                      ref_unet.diffusion_coeff_fn, cfg=cfgd, **kw)
            out[f"em_{name}_cfg/mean_x"] = emg.numpy()
        print("samplers", name, em.std().item(), pc.std().item())

    # --- DSM loss + gradients, 32x32 -------------------------------------------------------------
    cfg = config_for(n_lr=2, geo=True, seasons=True)
    b = synth_batch(batch=4, size=32, n_lr=2, geo=True, seasons=True)
    for mode in ("train", "eval"):
        net = build_reference(cfg)
        net.train(mode == "train")
        q = [uniform_draw(philox_ref.DRAW_DSM_T, (4,)), normal_draw(philox_ref.DRAW_DSM_Z, tuple(b.x.shape))]
        with injected_noise(q):
            loss = ref_unet.loss_fn(net, b.x, ref_unet.marginal_prob_std_fn, y=b.y, cond_img=b.cond_img,
                                    lsm_cond=b.lsm_cond, topo_cond=b.topo_cond, sdf_cond=b.sdf_cond)
        loss.backward()
        out[f"dsm_{mode}/loss"] = np.array(loss.item(), dtype=np.float64)
        gkeys = ["encoder.conv1.weight", "encoder.layer1.0.conv1.weight", "encoder.layer4.1.conv2.weight",
                 "encoder.attention_layers.3.mha.in_proj_weight", "decoder.residual_layers.0.conv_up.weight",
                 "decoder.residual_layers.3.norm2.weight", "decoder.final_layer.conv.weight",
                 "encoder.label_emb.weight", "encoder.time_projection_layers.2.1.weight"]
        params = dict(net.named_parameters())
        out[f"dsm_{mode}/grad_norms"] = np.array([params[k].grad.norm().item() for k in gkeys], dtype=np.float64)
        out[f"dsm_{mode}/grad_final_conv"] = params["decoder.final_layer.conv.weight"].grad.numpy()
        print("dsm", mode, loss.item())
    with open(os.path.join(HERE, "dsm_grad_keys.json"), "w") as f:
        json.dump(gkeys, f)

    # scalar schedules
    tt = torch.tensor([1e-3, 0.01, 0.1, 0.5, 0.9, 1.0])
    out["sde/t"] = tt.numpy()
    out["sde/std"] = ref_unet.marginal_prob_std_fn(tt).numpy()
    out["sde/g"] = ref_unet.diffusion_coeff_fn(tt).numpy()

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    with open(os.path.join(HERE, "reference_schema.json"), "w") as f:
        json.dump(schema, f)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
