"""CPU-only, world_size 2 over gloo: the data-parallel gradient exchange of sbgm_danra_b200.parallel
(DSM training).  The flat gradient buffer is all-reduced in place in contiguous
buckets that complete from the END of the buffer (backward order); the result must be the rank average
for every touched parameter regardless of bucket size, and untouched tail space must stay zero."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sbgm_danra_b200.parallel import GradBucketer, GradSync


def _layout():
    sizes = [("enc.a", 100), ("enc.b", 37), ("enc.unused", 50), ("dec.a", 300), ("dec.b", 12), ("tail.tp", 64)]
    layout, off = [], 0
    for name, n in sizes:
        layout.append((name, off, n))
        off += (n + 63) // 64 * 64
    return layout, off


def test_bucketer_partitions_and_completion_order():
    layout, total = _layout()
    b = GradBucketer(layout, total, bucket_elems=128, expected=[n for n, _, _ in layout if n != "enc.unused"])
    assert b.bounds[0][0] == 0 and b.bounds[-1][1] == total
    assert all(b.bounds[i][1] == b.bounds[i + 1][0] for i in range(len(b.bounds) - 1))
    # backward order: decoder first; a bucket fires exactly when its last expected gradient arrives
    fired = []
    for name in ["dec.b", "dec.a", "enc.b", "enc.a", "tail.tp"]:
        fired += b.touch([name])
    assert sorted(fired) == sorted(set(fired))
    assert set(fired) | set(b.remaining()) == set(range(len(b.bounds)))
    assert b.bucket_of["tail.tp"] in fired and b.bucket_of["dec.a"] in fired
    # unknown expected set (first step): nothing fires early, everything remains
    b0 = GradBucketer(layout, total, bucket_elems=128, expected=None)
    assert b0.touch(["dec.a", "enc.a"]) == [] and b0.remaining() == list(range(len(b0.bounds)))


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    layout, total = _layout()
    order = ["dec.b", "dec.a", "enc.b", "enc.a", "tail.tp"]       # "enc.unused" never receives a gradient
    ok = True
    for bucket_bytes in (64 * 4, 128 * 4, 1 << 20):
        sync = GradSync(None, bucket_bytes)
        for step in range(3):                                       # step 0 learns the expected set, 1.. overlap
            flat = torch.zeros(total)
            sync.begin(flat, layout)
            for name in order:
                off, n = next((o, k) for nm, o, k in layout if nm == name)
                flat[off:off + n] = torch.arange(n, dtype=torch.float32) * (rank + 1) + step
                sync.progress([name])
            sync.finish()
            for name, off, n in layout:
                want = torch.zeros(n) if name == "enc.unused" else torch.arange(n, dtype=torch.float32) * 1.5 + step
                ok = ok and torch.allclose(flat[off:off + n], want)
        ok = ok and sync.stats["overlapped"] > 0 if bucket_bytes < (1 << 20) else ok
    # a gradient that the first step did not produce ("enc.unused" appears from step 2 on, e.g. labels switched on): its bucket
    # has already been exchanged when it is touched -> finish() exchanges the slice on its own and the expected set grows
    sync = GradSync(None, 64 * 4)
    for step in range(4):
        flat = torch.zeros(total)
        sync.begin(flat, layout)
        for name in order + (["enc.unused"] if step >= 2 else []):
            off, n = next((o, k) for nm, o, k in layout if nm == name)
            flat[off:off + n] = torch.arange(n, dtype=torch.float32) * (rank + 1) + step
            sync.progress([name])
        sync.finish()
        for name, off, n in layout:
            want = torch.zeros(n) if (name == "enc.unused" and step < 2) else torch.arange(n, dtype=torch.float32) * 1.5 + step
            ok = ok and torch.allclose(flat[off:off + n], want)
    ok = ok and "enc.unused" in sync.expected
    # bf16 exchange: the average of the bf16-rounded rank gradients
    sync = GradSync(None, 1 << 20, grad_dtype="bf16")
    flat = torch.full((total,), 1.0 + rank)
    sync.begin(flat, layout)
    sync.progress(order)
    sync.finish()
    ok = ok and torch.allclose(flat, torch.full((total,), 1.5))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_bucketed_allreduce_averages_gradients():
    world = 2
    from conftest import free_port
    port = free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
