"""N>1 host logic on CPU with the gloo backend, world_size 2: batch sharding by independent samples, the flat-buffer
gradient all-reduce (equal to the single-process gradient of the concatenated batch), result gathering."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from magnet_b200 import distributed as D
from magnet_b200 import synthetic as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
        frozen = torch.nn.Linear(3, 3)           # a parameter that never receives a gradient on any rank
        batch = S.graph_batch(B=6, N=5, nt=6, d=2, seed=3)
        local = D.shard_batch(batch, rank, world)
        x = local["u"].reshape(-1, 6)
        loss = model(x).pow(2).mean()            # equal shard sizes: mean of means == global mean
        loss.backward()
        n = D.allreduce_gradients(list(model.parameters()) + list(frozen.parameters()), world)
        assert n == sum(p.numel() for p in model.parameters()) + sum(p.numel() for p in frozen.parameters())
        per_sample = local["u"].reshape(local["u"].shape[0], -1).sum(1, keepdim=True)
        sizes = [D.shard_range(6, r, world)[1] - D.shard_range(6, r, world)[0] for r in range(world)]
        gathered = D.gather_samples(per_sample, sizes)
        # the flat-buffer path of optim.FlatAdam: one collective on the buffer itself, the mean applied as grad_scale
        from magnet_b200.optim import allreduce_flat_gradient

        class _Flat:          # stands in for FlatAdam (which needs CUDA parameters): only .flat_grad is touched
            flat_grad = torch.full((5,), float(rank + 1))
        scale = allreduce_flat_gradient(_Flat, world)
        torch.save({"grads": [p.grad.clone() for p in model.parameters()], "frozen": [p.grad.clone() for p in frozen.parameters()],
                    "gathered": gathered, "lo_hi": D.shard_range(6, rank, world), "flat": _Flat.flat_grad.clone(), "scale": scale},
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 64):
        for w in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "r0.pt")
    r1 = torch.load(tmp_path / "r1.pt")
    assert r0["lo_hi"] == (0, 3) and r1["lo_hi"] == (3, 6)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    batch = S.graph_batch(B=6, N=5, nt=6, d=2, seed=3)
    model(batch["u"].reshape(-1, 6)).pow(2).mean().backward()
    for g0, g1, p in zip(r0["grads"], r1["grads"], model.parameters()):
        assert torch.equal(g0, g1)                                   # both ranks hold the same reduced gradient
        assert torch.allclose(g0, p.grad, rtol=1e-6, atol=1e-8)      # and it is the full-batch gradient
    assert all(float(g.abs().max()) == 0.0 for g in r0["frozen"])
    want = batch["u"].reshape(6, -1).sum(1, keepdim=True)
    assert torch.equal(r0["gathered"], want) and torch.equal(r1["gathered"], want)
    assert torch.equal(r0["flat"], torch.full((5,), 3.0)) and torch.equal(r1["flat"], r0["flat"]) and r0["scale"] == 0.5


def _overlap_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from magnet_b200.optim import OverlappedFlatAllReduce
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
        params = list(model.parameters())

        class _Flat:          # the part of FlatAdam the all-reduce touches (FlatAdam itself needs CUDA parameters)
            pass
        opt, offs, n = _Flat(), [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        opt._params, opt._offs, opt.flat_grad = params, offs, torch.zeros(n)
        for p, o in zip(params, offs):
            p.grad = opt.flat_grad[o:o + p.numel()].view_as(p)
        ar = OverlappedFlatAllReduce(opt, world, n_buckets=3)
        batch = S.graph_batch(B=6, N=5, nt=6, d=2, seed=3)
        local = D.shard_batch(batch, rank, world)
        launched_during_backward = []
        for step in range(2):                      # twice: the bucket counters reset
            opt.flat_grad.zero_()
            model(local["u"].reshape(-1, 6)).pow(2).mean().backward()
            launched_during_backward.append(list(ar._launched))
            scale = ar.finish()
        torch.save({"flat": opt.flat_grad.clone() * scale, "buckets": ar.buckets, "launched": launched_during_backward},
                   os.path.join(out_dir, f"o{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_two_rank_overlapped_bucketed_allreduce(tmp_path):
    """OverlappedFlatAllReduce: every bucket is launched from the gradient hooks during backward, the buckets tile the flat
    buffer, and the result is the full-batch gradient on both ranks."""
    world = 2
    mp.spawn(_overlap_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = torch.load(tmp_path / "o0.pt"), torch.load(tmp_path / "o1.pt")
    assert torch.equal(r0["flat"], r1["flat"])
    assert all(all(l) for l in r0["launched"]) and len(r0["launched"]) == 2          # all three buckets left during backward
    b = r0["buckets"]
    assert b[0][2] == 0 and all(x[3] == y[2] for x, y in zip(b, b[1:])) and b[-1][3] == r0["flat"].numel()
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
    batch = S.graph_batch(B=6, N=5, nt=6, d=2, seed=3)
    model(batch["u"].reshape(-1, 6)).pow(2).mean().backward()
    off = 0
    for p in model.parameters():
        assert torch.allclose(r0["flat"][off:off + p.numel()].view_as(p), p.grad, rtol=1e-6, atol=1e-8)
        off += (p.numel() + 3) // 4 * 4
