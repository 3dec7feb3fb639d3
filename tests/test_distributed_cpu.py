"""world_size-2 gloo test of the N>1 path: frames shard with no data-path collective and the
gathered per-rank results equal the single-process result (host logic only, CPU)."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from pdfnet_b200 import parallel, synth
    r, w, _ = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    lo, hi = parallel.shard_range(n_frames, r, w)
    clouds = synth.clouds(n_frames, n_points=64, seed=9)          # every rank can build the full batch
    local = clouds[lo:hi].sum(dim=1)                                # stand-in for the per-frame hot path
    full = parallel.gather_results(local, n_frames)
    t = parallel.max_over_ranks(float(rank + 1), "cpu")
    torch.save(dict(full=full, lo=lo, hi=hi, t=t), os.path.join(out_dir, "r%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding(tmp_path):
    n_frames, world = 11, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_frames, str(tmp_path)), nprocs=world, join=True)
    from pdfnet_b200 import synth
    ref = synth.clouds(n_frames, n_points=64, seed=9).sum(dim=1)
    res = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    assert res[0]["lo"] == 0 and res[0]["hi"] == res[1]["lo"] and res[1]["hi"] == n_frames
    for r in res:
        assert torch.equal(r["full"], ref)
        assert r["t"] == 2.0


def _grad_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from pdfnet_b200 import parallel, training
    parallel.init_distributed("gloo")
    g = torch.Generator().manual_seed(100 + rank)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((3, 5), (7,), (2, 2, 2))]
    params.append(torch.nn.Parameter(torch.zeros(4)))               # no gradient: skipped, like unused heads
    for p in params[:3]:
        p.grad = torch.randn(p.shape, generator=g)
    nbytes = training.allreduce_gradients(params, world)
    torch.save(dict(grads=[p.grad for p in params[:3]], nbytes=nbytes, none=params[3].grad is None),
               os.path.join(out_dir, "g%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce(tmp_path):
    """The cfg5 exchange step (DDP semantics, main.py:44-73): every rank ends with the MEAN gradient."""
    world = 2
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "g%d.pt" % r)) for r in range(world)]
    shapes = ((3, 5), (7,), (2, 2, 2))
    expect = []
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    for s in shapes:
        expect.append(sum(torch.randn(s, generator=g) for g in gens) / world)
    for r in res:
        assert r["none"] and r["nbytes"] == 4 * (15 + 7 + 8)
        for got, want in zip(r["grads"], expect):
            torch.testing.assert_close(got, want)


def _bucket_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from pdfnet_b200 import parallel, training
    parallel.init_distributed("gloo")
    torch.manual_seed(5)                                           # same weights on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                              torch.nn.Linear(16, 3))
    unused = torch.nn.Parameter(torch.zeros(5))                    # never receives a gradient (unused head)
    params = list(net.parameters()) + [unused]
    sync = training.BucketedAllReduce(params, world, bucket_bytes=512)     # tiny buckets: several per step
    x = torch.randn((8, 6), generator=torch.Generator().manual_seed(200 + rank))
    outs = []
    for step in range(2):                                          # second step: buckets re-armed, views intact
        sync.zero_grad()
        net(x * (step + 1)).pow(2).sum().backward()
        nbytes = sync.finish()
        outs.append([p.grad.clone() for p in params])
    views = all(p.grad.data_ptr() >= b["flat"].data_ptr() for b in sync.buckets for p in b["params"])
    torch.save(dict(outs=outs, nbytes=nbytes, n_buckets=len(sync.buckets), views=views),
               os.path.join(out_dir, "b%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucketed_overlapped_allreduce(tmp_path):
    """BucketedAllReduce (the cfg5 exchange, DDP semantics of base_trainer.py:94-95): gradients are views of
    flat buckets, each bucket is all-reduced from a gradient hook as soon as it is complete, every rank ends
    with the MEAN gradient, parameters without a gradient exchange zeros."""
    world = 2
    mp.spawn(_bucket_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "b%d.pt" % r)) for r in range(world)]
    torch.manual_seed(5)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                              torch.nn.Linear(16, 3))
    for step in range(2):
        want = None
        for r in range(world):
            net.zero_grad()
            x = torch.randn((8, 6), generator=torch.Generator().manual_seed(200 + r))
            net(x * (step + 1)).pow(2).sum().backward()
            g = [p.grad.clone() for p in net.parameters()]
            want = g if want is None else [a + b for a, b in zip(want, g)]
        want = [w / world for w in want]
        for r in res:
            for got, w in zip(r["outs"][step][:-1], want):
                torch.testing.assert_close(got, w)
            assert float(r["outs"][step][-1].abs().max()) == 0.0
    n_el = sum(p.numel() for p in net.parameters()) + 5
    assert res[0]["n_buckets"] >= 2 and res[0]["views"] and res[0]["nbytes"] == 4 * n_el
