"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py

Every rank samples its shard of a small ensemble with the EM sampler (no collective) and the PC sampler (one
all-gather of per-member gradient norms per step, captured in the CUDA graph); rank 0 also samples the whole
ensemble alone and checks that the gathered shards reproduce it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
from sbgm_danra_b200 import score_sampling as ss
from sbgm_danra_b200._smoke import build_model
from sbgm_danra_b200.score_unet import diffusion_coeff_fn, marginal_prob_std_fn


def train_check(rank, world, dev, net, cfg, ck) -> bool:
    """Data-parallel DSM step: every rank takes its slice of one global batch, gradients are averaged by the bucketed
    all-reduce (sbgm_danra_b200.parallel); rank 0 recomputes the full batch alone.  BatchNorm in eval mode (constant
    statistics) so that the per-sample losses do not couple and the two must agree to rounding.  Steps 3+ replay the
    captured CUDA graphs with the NCCL all-reduces inside."""
    from sbgm_danra_b200 import parallel
    from sbgm_danra_b200.score_unet import loss_fn
    per, size = 2, 32
    total = per * world
    b = synth_batch(batch=total, size=size, shared_cond=False, **ck)
    cut = lambda v, s: None if v is None else v[s].to(dev)

    def grads(sl, first, members):
        ss.manual_seed(77)
        ss.set_ensemble_shard(first, members, None)
        net.zero_grad(set_to_none=True)
        loss = loss_fn(net, cut(b.x, sl), marginal_prob_std_fn, y=cut(b.y, sl), cond_img=cut(b.cond_img, sl),
                       lsm_cond=cut(b.lsm_cond, sl), topo_cond=cut(b.topo_cond, sl), sdf_cond=cut(b.sdf_cond, sl))
        loss.backward()
        return loss.detach(), {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}

    ok = True
    for mode, sync_bn in (("eval-mode BatchNorm", False), ("train-mode BatchNorm, synchronised statistics", True)):
        net.train(sync_bn)
        sync = parallel.attach(net, bucket_bytes=4 << 20, sync_bn=sync_bn)
        for step in range(4):
            loss, g = grads(slice(rank * per, (rank + 1) * per), rank * per, total)
            dist.all_reduce(loss)
            loss /= world
        parallel.detach(net)
        net._train_runners.clear()
        if rank == 0:
            loss_full, g_full = grads(slice(None), 0, None)
            num = sum(float((g[k] - g_full[k]).double().pow(2).sum()) for k in g_full)
            den = sum(float(g_full[k].double().pow(2).sum()) for k in g_full)
            err = (num / den) ** 0.5
            print(f"[multi-gpu x{world}] DSM step ({mode}): averaged sharded gradients vs single-GPU full batch rel-L2 = {err:.3e}; "
                  f"loss {float(loss):.6f} vs {float(loss_full):.6f}; buckets {sync.stats}")
            ok = ok and err < 1e-4 and abs(float(loss) - float(loss_full)) / abs(float(loss_full)) < 1e-5 and sync.stats["overlapped"] > 0
        dist.barrier()
        net._train_runners.clear()
    net.eval()
    return ok


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ck = dict(n_lr=2, geo=True, seasons=True)
    cfg = config_for(**ck)
    net = build_model(cfg, synth_state_dict(cfg), "bf16x3", dev)
    per, size, steps = 2, 32, 5
    total = per * world
    b = synth_batch(batch=total, size=size, shared_cond=False, **ck)
    sl = slice(rank * per, (rank + 1) * per)

    def cut(v, s):
        return None if v is None else v[s].to(dev)

    ok = True
    for name, fn in (("em", ss.Euler_Maruyama_sampler), ("pc", ss.pc_sampler)):
        ss.manual_seed(123)
        ss.set_ensemble_shard(rank * per, total, None)
        part = fn(net, marginal_prob_std_fn, diffusion_coeff_fn, batch_size=per, num_steps=steps, device=dev, img_size=size,
                  y=cut(b.y, sl), cond_img=cut(b.cond_img, sl), lsm_cond=cut(b.lsm_cond, sl), topo_cond=cut(b.topo_cond, sl))
        gathered = torch.empty((total, 1, size, size), device=dev)
        dist.all_gather_into_tensor(gathered, part.contiguous())
        if rank == 0:
            ss.manual_seed(123)
            ss.set_ensemble_shard(0, None, None)
            full = fn(net, marginal_prob_std_fn, diffusion_coeff_fn, batch_size=total, num_steps=steps, device=dev,
                      img_size=size, y=cut(b.y, slice(None)), cond_img=cut(b.cond_img, slice(None)),
                      lsm_cond=cut(b.lsm_cond, slice(None)), topo_cond=cut(b.topo_cond, slice(None)))
            err = float((gathered - full).norm() / full.norm())
            print(f"[multi-gpu x{world}] {name}: sharded vs single-GPU rel-L2 = {err:.3e}")
            ok = ok and err < 1e-4
        dist.barrier()
    # train-mode BatchNorm while sampling (generation.py:47 quirk): the members are coupled through the batch statistics, so a
    # sharded ensemble all-gathers the per-channel partial sums inside the captured step (synchronised BatchNorm); the gathered
    # shards must reproduce rank 0 sampling all members alone, and the running statistics must move identically
    net.train()
    state0 = {k: v.clone() for k, v in net.state_dict().items()}
    for name, fn in (("em", ss.Euler_Maruyama_sampler), ("pc", ss.pc_sampler)):
        net.load_state_dict(state0)
        ss.manual_seed(321)
        ss.set_ensemble_shard(rank * per, total, dist.group.WORLD)
        part = fn(net, marginal_prob_std_fn, diffusion_coeff_fn, batch_size=per, num_steps=steps, device=dev, img_size=size,
                  y=cut(b.y, sl), cond_img=cut(b.cond_img, sl), lsm_cond=cut(b.lsm_cond, sl), topo_cond=cut(b.topo_cond, sl))
        rm_sharded = net.encoder.layer3[0].bn1.running_mean.clone()
        gathered = torch.empty((total, 1, size, size), device=dev)
        dist.all_gather_into_tensor(gathered, part.contiguous())
        ss.clear_sampler_cache()
        if rank == 0:
            net.load_state_dict(state0)
            ss.manual_seed(321)
            ss.set_ensemble_shard(0, None, None)
            full = fn(net, marginal_prob_std_fn, diffusion_coeff_fn, batch_size=total, num_steps=steps, device=dev,
                      img_size=size, y=cut(b.y, slice(None)), cond_img=cut(b.cond_img, slice(None)),
                      lsm_cond=cut(b.lsm_cond, slice(None)), topo_cond=cut(b.topo_cond, slice(None)))
            err = float((gathered - full).norm() / full.norm())
            rm_err = float((rm_sharded - net.encoder.layer3[0].bn1.running_mean).norm() / net.encoder.layer3[0].bn1.running_mean.norm())
            print(f"[multi-gpu x{world}] {name}, train-mode BatchNorm with synchronised statistics: sharded vs single-GPU rel-L2 = {err:.3e}; "
                  f"running_mean rel {rm_err:.1e}")
            ok = ok and err < 1e-4 and rm_err < 1e-5
            ss.clear_sampler_cache()
        dist.barrier()
    net.load_state_dict(state0)
    net.eval()
    ss.set_ensemble_shard(0, None, None)
    ok = train_check(rank, world, dev, net, cfg, ck) and ok
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    good = int(flag.item()) == 1
    if rank == 0 and good:
        print("multi-gpu check OK", flush=True)
    # the cached sampler plans hold CUDA graphs with captured NCCL all-gathers: drop them before the
    # communicator goes away, and skip interpreter teardown (destroying a communicator that graphs still
    # reference can block)
    ss.clear_sampler_cache()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if good else 1)


if __name__ == "__main__":
    main()
