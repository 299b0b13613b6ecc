"""World-size-2 gloo tests (CPU) of the data-parallel host logic (mrn_b200/dist.py): the ONE exchange step of the
path -- an all-reduce (average) of the flat gradient arena -- plus sharding and broadcast helpers."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from mrn_b200 import dist as mdist
    r, lr, w = mdist.init_from_env("gloo")
    assert (r, w) == (rank, world) and mdist.world_size() == world
    # gradient arena: rank-dependent values -> mean over ranks, identical on every rank
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    mdist.allreduce_mean_(g)
    expect = torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(g, expect)
    # parameter broadcast from rank 0
    p = torch.full((17,), float(rank))
    mdist.broadcast_(p, src=0)
    ok = ok and bool((p == 0).all())
    lo, hi = mdist.shard_bounds(512, rank, world)
    ok = ok and (hi - lo == 512 // world) and lo == rank * (512 // world)
    mx = mdist.max_over_ranks(float(rank + 1), torch.device("cpu"))
    ok = ok and mx == float(world)
    mdist.barrier()
    out[rank] = ok
    torch.distributed.destroy_process_group()


def test_gloo_world2_allreduce_broadcast_shards():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_shard_bounds_rejects_unequal_shards():
    from mrn_b200 import dist as mdist
    with pytest.raises(ValueError):
        mdist.shard_bounds(10, 0, 3)


def test_mean_of_shard_means_equals_global_mean():
    """Equal shards make the per-rank CTC mean + gradient average exact (SURVEY.md §7 'DataParallel semantics')."""
    torch.manual_seed(0)
    nll = torch.rand(64, dtype=torch.float64)
    lens = torch.randint(1, 26, (64,)).double()
    full = (nll / lens).mean()
    halves = [(nll[i * 32:(i + 1) * 32] / lens[i * 32:(i + 1) * 32]).mean() for i in range(2)]
    assert abs(float(full) - float(sum(halves) / 2)) < 1e-12


def _worker_stage0(rank, world, port, out):
    """Each rank: oracle stage-0 gradients of its shard (eval-mode BatchNorm) flattened into one arena, averaged over gloo."""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from mrn_b200 import dist as mdist
    from oracle import mrn_oracle as O
    from oracle import synth
    mdist.init_from_env("gloo")
    cc, B, seed = (29,), 4, 41
    sd = synth.synth_state_dict(cc, seed, arch="crnn")
    img, tgt, lens, _ = synth.synth_batch(B, cc, seed)
    lo, hi = mdist.shard_bounds(B, rank, world)
    r = O.stage0_loss_and_grads(sd, 0, img[lo:hi], tgt[lo:hi], lens[lo:hi], "eval", None, dtype=torch.float64)
    keys = sorted(r["grads"])
    arena = torch.cat([r["grads"][k].reshape(-1) for k in keys] + [r["loss"].reshape(1)])      # the loss rides along
    mdist.allreduce_mean_(arena)
    out[rank] = arena
    torch.distributed.destroy_process_group()


def test_gloo_world2_stage0_gradient_arena_equals_the_full_batch():
    """The stage-0 exchange step (one all-reduce of the expert's flat gradient arena) on two gloo ranks:
    the averaged shard gradients and loss equal the full-batch oracle step (running-statistics BatchNorm; batch
    statistics stay per rank, as under nn.DataParallel)."""
    from oracle import mrn_oracle as O
    from oracle import synth
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_stage0, args=(world, port, out), nprocs=world, join=True)
    cc, B, seed = (29,), 4, 41
    sd = synth.synth_state_dict(cc, seed, arch="crnn")
    img, tgt, lens, _ = synth.synth_batch(B, cc, seed)
    r = O.stage0_loss_and_grads(sd, 0, img, tgt, lens, "eval", None, dtype=torch.float64)
    keys = sorted(r["grads"])
    ref = torch.cat([r["grads"][k].reshape(-1) for k in keys] + [r["loss"].reshape(1)])
    assert torch.equal(out[0], out[1])
    assert float((out[0] - ref).abs().max()) < 1e-9 * max(1.0, float(ref.abs().max()))
