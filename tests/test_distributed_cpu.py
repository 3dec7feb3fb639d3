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
