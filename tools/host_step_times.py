"""Host-side time of each phase of a DSM training step (C4 shape) with the GPU idle at the start of the step, as in the
reference loop (sbgm/training.py:410 reads `batch_loss.item()` every step): how long the GPU waits for its first launch.

    python tools/host_step_times.py [--steps 20]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    from sbgm_danra_b200 import optim as sbgm_optim, score_sampling
    from sbgm_danra_b200._smoke import build_model
    from sbgm_danra_b200.score_unet import loss_fn, marginal_prob_std_fn
    dev = "cuda:0"
    cfg = config_for(n_lr=2, geo=True, seasons=True)
    net = build_model(cfg, synth_state_dict(cfg), a.precision, dev).train()
    b = synth_batch(batch=64, size=128, n_lr=2, geo=True, seasons=True, seed=1234)
    c = lambda v: None if v is None else v.to(dev)
    x, y, cond, lsm, topo, sdf = c(b.x), c(b.y), c(b.cond_img), c(b.lsm_cond), c(b.topo_cond), c(b.sdf_cond)
    opt = sbgm_optim.Adam(net.parameters(), lr=1e-4)
    score_sampling.manual_seed(5)
    acc = [0.0] * 5
    for it in range(a.steps + 5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        t1 = time.perf_counter()
        loss = loss_fn(net, x, marginal_prob_std_fn, y=y, cond_img=cond, lsm_cond=lsm, topo_cond=topo, sdf_cond=sdf)
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        opt.step()
        t4 = time.perf_counter()
        v = loss.item()
        t5 = time.perf_counter()
        if it >= 5:
            for k, d in enumerate((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t0)):
                acc[k] += d
    n = a.steps
    print(f"host ms per step: zero_grad {acc[0] / n * 1e3:.3f}  loss_fn {acc[1] / n * 1e3:.3f}  backward {acc[2] / n * 1e3:.3f}  "
          f"optimizer.step {acc[3] / n * 1e3:.3f}  whole step incl. loss.item() {acc[4] / n * 1e3:.3f}  (loss {v:.1f})")


if __name__ == "__main__":
    main()
