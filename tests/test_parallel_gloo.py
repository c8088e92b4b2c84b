"""world_size-2/3 gloo runs of the multi-GPU host logic on CPU (the N>1 path of DESIGN.md 'Multi-GPU')."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dmvsnet_b200 import parallel


def test_plane_shards_cover_exactly():
    for d in (48, 32, 8, 4, 5, 1):
        for world in (1, 2, 3, 4, 8):
            sh = parallel.plane_shards(d, world)
            assert len(sh) == world and sh[0][0] == 0 and sh[-1][1] == d
            assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
            sizes = [b - a for a, b in sh]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.view_shard(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((parallel.view_shard(11, r, 8) for r in range(8)), [])) == list(range(11))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, planes, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from dmvsnet_b200 import synthetic as syn
        from oracle import dmvs_oracle as O
        g = torch.Generator().manual_seed(3)
        b, c, h, w, n = 2, 8, 12, 20, 3
        feats = [torch.randn(b, c, h, w, generator=g) for _ in range(n)]
        proj = syn.make_proj_matrices(h * 4, w * 4, n, b, num_stages=1)["stage1"]
        hyp = 425 + 500 * torch.rand(b, planes, h, w, generator=g)
        full = O.warp_corr(feats, proj, hyp)  # every rank knows the unsharded answer

        def compute(lo, hi):  # the CPU stand-in for ops.warp_corr(..., d_range=(lo, hi))
            return O.warp_corr(feats, proj, hyp[:, lo:hi].contiguous())

        got = parallel.gather_planes(compute, planes, plane_dim=2)
        ok = torch.equal(got, full)
        open(os.path.join(result_dir, "rank%d" % rank), "w").write("ok" if ok else "mismatch %g" % float((got - full).abs().max()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,planes", [(2, 8), (2, 5), (3, 4), (3, 2)])
def test_depth_sharded_cost_volume_equals_unsharded(tmp_path, world, planes):
    """Plane shards computed independently + one all-gather == the unsharded cost volume, bit for bit (also with
    uneven shards and with more ranks than planes)."""
    port = _free_port()
    mp.spawn(_worker, args=(world, port, planes, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"


def _branch_worker(rank, world, port, batch, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11)
        full = torch.randn(batch, 4, 3, 5, 7, generator=g)       # what an unsharded regnet would return (same on every rank)
        mine = torch.full_like(full, float("nan"))                # rank r only computes its branch: channels 2r, 2r+1
        mine[:, 2 * rank:2 * rank + 2] = full[:, 2 * rank:2 * rank + 2]
        got = parallel.exchange_channel_halves(mine)
        open(os.path.join(result_dir, "rank%d" % rank), "w").write("ok" if torch.equal(got, full) else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [1, 2])
def test_branch_sharded_logits_exchange(tmp_path, batch):
    """Regularisation-net branch sharding over 2 ranks: each rank fills its two logit channels, one all-gather completes both."""
    port = _free_port()
    mp.spawn(_branch_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"
