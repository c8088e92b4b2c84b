"""torchrun --nproc-per-node N tools/check_depth_shard_nccl.py - W1 split by depth plane over N GPUs + one NCCL all-gather
(parallel.warp_corr_depth_sharded) against the unsharded kernel on every rank: must be bit-identical.  Also times both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from dmvsnet_b200 import ops, parallel, synthetic as syn

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
saved = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=dev); dist.barrier(); torch.cuda.synchronize()
os.dup2(saved, 1); os.close(saved)
g = torch.Generator().manual_seed(0)  # same inputs on every rank (features / hypotheses are replicated)
ok = True
for (c, d, h, w, n) in ((32, 48, 296, 400, 5), (16, 32, 592, 800, 5), (8, 8, 1184, 1600, 5), (8, 4, 296, 400, 11)):
    feats = [torch.randn(1, c, h, w, generator=g).to(dev) for _ in range(n)]
    rt = ops.relative_projections(syn.make_proj_matrices(h * 4, w * 4, n, 1, num_stages=1)["stage1"]).to(dev)
    hyp = (425 + 500 * torch.rand(1, d, h, w, generator=g)).to(dev)
    for layout in ("nhwc", "staged"):
        ops.W1_LAYOUT = layout
        full = ops.warp_corr(feats, rt, hyp)
        shard = parallel.warp_corr_depth_sharded(feats, rt, hyp)
        same = bool(torch.equal(full, shard))
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())
        def timeit(fn):
            for _ in range(2): fn()
            dist.barrier(); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): fn()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        t_full = timeit(lambda: ops.warp_corr(feats, rt, hyp))
        t_shard = timeit(lambda: parallel.warp_corr_depth_sharded(feats, rt, hyp))
        if rank == 0:
            print("C=%d D=%d %dx%d N=%d %-6s world=%d  identical=%s  unsharded %.3f ms  sharded+all-gather %.3f ms" %
                  (c, d, h, w, n, layout, world, bool(flag.item()), t_full, t_shard), flush=True)
if rank == 0:
    print("ALL IDENTICAL" if ok else "MISMATCH")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
