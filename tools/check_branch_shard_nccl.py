"""torchrun --nproc-per-node 2 tools/check_branch_shard_nccl.py - single-view latency mode on 2 GPUs: every regularisation net runs
one branch per rank + one NCCL all-gather of the logit halves (MVSNet.cascade(branch_group=...)).  Must reproduce the single-GPU
cascade bit for bit; prints both hot-path times (max over ranks, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from dmvsnet_b200 import MVSNet, synthetic as syn

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
assert world == 2
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
saved = os.dup(1); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=dev); dist.barrier(); torch.cuda.synchronize()
os.dup2(saved, 1); os.close(saved)
ok = True
for (H, W, views, nd) in ((128, 160, 3, [16, 8, 8]), (1184, 1600, 5, [48, 32, 8])):
    ratios = [4, 2, 1]
    net = MVSNet(nd, ratios, inverse_depth=True)
    net.load_state_dict(syn.randomise_regnet_state(net.state_dict(), seed=0))
    net = net.to(dev).eval()
    imgs = syn.make_images(H, W, views, 1, seed=0, natural=True).to(dev)  # same view set on both ranks
    proj = syn.make_proj_matrices(H, W, views, 1, num_stages=3)
    dv = syn.make_depth_values(1, 192, inverse=True).to(dev)
    with torch.no_grad():
        feats = net.extract_features(imgs)
        single = net.cascade(feats, proj, dv, (H, W))
        shard = net.cascade(feats, proj, dv, (H, W), branch_group=dist.group.WORLD)
        same = all(bool(torch.equal(single[k], shard[k])) for k in ("depth", "photometric_confidence", "prob_volume", "depth_values_c"))
        flag = torch.tensor([int(same)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = ok and bool(flag.item())

        def timeit(fn, n=5):
            for _ in range(3): fn()
            dist.barrier(); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        t1 = timeit(lambda: net.cascade(feats, proj, dv, (H, W)))
        t2 = timeit(lambda: net.cascade(feats, proj, dv, (H, W), branch_group=dist.group.WORLD))
    if rank == 0:
        print("%dx%d N=%d D=%s: identical=%s  one GPU %.2f ms  two GPUs (branch-sharded regnets) %.2f ms  -> x%.2f" %
              (W, H, views, nd, bool(flag.item()), t1, t2, t1 / t2), flush=True)
if rank == 0:
    print("ALL IDENTICAL" if ok else "MISMATCH")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
