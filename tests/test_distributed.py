"""CPU, gloo, world_size 2: the multi-GPU path shards batch rows across ranks with NO data-path collective (sequences are
independent units).  Checks the sharding helper and that gathering the per-rank results reproduces the
unsharded forward (emulated ops stand in for the kernels; the arithmetic per row is identical by construction)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import emu_ops
    import vpt_b200
    from common import make_policy, small_kwargs
    from video_pre_training_b200 import ops, parallel

    for name in dir(emu_ops):
        if not name.startswith("_") and callable(getattr(emu_ops, name)) and hasattr(ops, name):
            setattr(ops, name, getattr(emu_ops, name))
    pol, sd, cfg = make_policy(small_kwargs(), seed=0)
    B, T = 5, 8  # uneven split on purpose: ranks get 3 and 2 rows
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (B, T, 32, 32, 3), dtype=torch.uint8, generator=g)
    first = torch.zeros(B, T, dtype=torch.bool)
    lo, hi = parallel.shard_range(B, rank, world)
    (pd, v, _), st = pol({"img": img[lo:hi]}, first[lo:hi], pol.initial_state(hi - lo))
    full = parallel.all_gather_rows(pd["camera"], B)
    if rank == 0:
        (pd_all, _, _), _ = pol({"img": img}, first, pol.initial_state(B))
        q.put((lo, hi, bool(torch.equal(full, pd_all["camera"])), tuple(full.shape)))
    else:
        q.put((lo, hi, True, tuple(full.shape)))
    dist.destroy_process_group()


def test_batch_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=240) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[:2] for r in res] == [(0, 3), (3, 5)]
    assert all(r[2] for r in res) and all(r[3] == (5, 8, 1, 121) for r in res)


def _adam_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vpt_b200  # noqa: F401
    from video_pre_training_b200.parallel import FlatAdamDP

    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    opt = FlatAdamDP(params, lr=1e-3)
    ok_alias = params[0].data_ptr() == opt.flat_p.data_ptr() and params[1].grad.data_ptr() == opt.flat_g[16:].data_ptr()
    opt.zero_grad()
    params[0].grad.fill_(float(rank + 1))     # rank-dependent gradients written through the aliased views
    params[1].grad.fill_(10.0 * (rank + 1))
    w = opt.reduce_gradients()                # ONE all-reduce over the flat bucket
    q.put((rank, w, ok_alias, params[0].grad.flatten()[0].item(), params[1].grad.flatten()[0].item()))
    dist.destroy_process_group()


def test_flat_gradient_bucket_single_allreduce_gloo():
    """BC data-parallel plumbing: parameters / gradients alias flat buckets; one sum all-reduce."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_adam_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, w, ok_alias, g0, g1 in res:
        assert w == 2 and ok_alias and g0 == 3.0 and g1 == 30.0


def test_shard_range_covers_everything():
    import vpt_b200  # noqa: F401  (registers the hyphenated package directory as video_pre_training_b200)
    from video_pre_training_b200 import parallel

    for B in (1, 5, 128):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _bc_worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import vpt_b200  # noqa: F401
    from common import emulation, make_policy, small_kwargs
    from video_pre_training_b200 import parallel
    from video_pre_training_b200.training import BCTrainer

    pol, _, _ = make_policy(small_kwargs(), seed=0)  # identical replicas
    B, T = 4, 8
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (B, T, 32, 32, 3), dtype=torch.uint8, generator=g)
    first = torch.zeros(B, T, dtype=torch.bool)
    actions = {"camera": torch.randint(0, 121, (B, T, 1), generator=g), "buttons": torch.randint(0, 8641, (B, T, 1), generator=g)}
    named = [(n, p) for n, p in pol.named_parameters() if not n.startswith("value_head")]
    opt = parallel.FlatAdamDP([p for _, p in named], lr=1e-3)
    lo, hi = parallel.shard_range(B, rank, world)
    with emulation():
        opt.zero_grad()
        shard = (img[lo:hi], first[lo:hi], pol.initial_state(hi - lo), {k: v[lo:hi] for k, v in actions.items()})
        BCTrainer(pol).loss_and_grad(*shard)
        w = opt.reduce_gradients()                      # the step's single collective
        dp_grad = opt.flat_g.clone() / w                # (the 1/world lives in the Adam kernel)
        # overlapped variant: the slice from the dense layer on is reduced while the ImpalaCNN backward still runs
        split = opt.offset_of(pol.net.img_process.cnn.dense.norm.weight)
        assert 0 < split < opt.n and opt.offset_of(named[0][1]) == 0
        opt.zero_grad()
        BCTrainer(pol).loss_and_grad(*shard, upper_grads_ready=lambda: opt.reduce_async(split, opt.n))
        assert opt._pending is not None
        opt.reduce_gradients()
        assert opt._pending is None and torch.equal(opt.flat_g / w, dp_grad)
        if rank == 0:                                   # the same global batch in one process
            opt.zero_grad()
            BCTrainer(pol).loss_and_grad(img, first, pol.initial_state(B), actions)
            err = ((dp_grad - opt.flat_g).norm() / opt.flat_g.norm()).item()
            q.put((rank, w, err))
        else:
            q.put((rank, w, 0.0))
    dist.destroy_process_group()


def test_bc_data_parallel_gradients_equal_the_global_batch_gloo():
    """BC step over 2 ranks: clips sharded across ranks, replicated weights, ONE all-reduce of the flat gradient
    bucket; the averaged result equals the gradient of the whole batch computed in one process (sequences are independent, the
    loss is a mean over frames; bf16 rounding of per-rank partial sums is the only difference)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_bc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=600) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(w == 2 for _, w, _ in res)
    assert res[0][2] < 2e-2, res
