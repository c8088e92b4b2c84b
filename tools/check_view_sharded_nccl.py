"""Single-view sharded mode under torchrun (N ranks, NCCL): bit identity against the one-GPU forward and a phase breakdown.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_view_sharded_nccl.py [--config tnt]

Per phase (rank 0, CUDA events on the compute stream; the host column is the time the Python loop needs to enqueue the phase -
when it is close to the device column the phase is launch bound): FeatureNet of this rank's views + issue of the six fp16
all-gathers, then the three stages of the row-band cascade (each including its waits on outstanding gathers and its two
all-reduces).  Output is kept under profiles/.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from dmvsnet_b200 import parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="tnt")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    wl = bench.make_workload(args.config, seed=0, want_features=False)
    net = wl["net"]
    net.load_state_dict(wl["state"])
    net = net.to(dev).eval()
    net.DepthNet.return_prob_volume = False
    imgs, proj, dv = wl["imgs"].to(dev), wl["proj"], wl["dv"].to(dev)
    with torch.no_grad():
        ref = net(imgs, proj, dv)
        for _ in range(3):
            out = parallel.infer_view_sharded(net, imgs, proj, dv)
        same = all(torch.equal(out["stage%d" % s][k], ref["stage%d" % s][k]) for s in (1, 2, 3)
                   for k in ("depth", "photometric_confidence", "photometric_confidence_refine", "depth_values_c", "depth_sub_plus"))
        # phase breakdown: re-implement infer_view_sharded with events between the phases
        rows = []
        for _ in range(args.steps):
            dist.barrier(); torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            t = [time.perf_counter()]
            ev[0].record()
            r, s_, wait_for = parallel.extract_features_view_sharded(net, imgs)
            ev[1].record(); t.append(time.perf_counter())
            marks = []
            o = parallel.cascade_row_sharded(net, r, s_, proj, dv, imgs.shape[-2:], parallel.exchange_group(), wait_for=wait_for, marks=marks)
            ev[2].record(); t.append(time.perf_counter())
            torch.cuda.synchronize(); t.append(time.perf_counter())
            phases = [(b[0], a[1].elapsed_time(b[1])) for a, b in zip(marks, marks[1:])]
            rows.append([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), (t[1] - t[0]) * 1e3, (t[2] - t[1]) * 1e3, (t[3] - t[0]) * 1e3])
        med = [sorted(c)[len(c) // 2] for c in zip(*rows)]
    res = torch.tensor([0.0 if same else 1.0] + med, device=dev, dtype=torch.float64)
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": args.config, "ranks": world, "bit_identical_all_maps": bool(res[0] == 0),
                          "device_ms": {"featurenet_and_gather_issue": float(res[1]), "cascade": float(res[2])},
                          "host_enqueue_ms": {"featurenet_and_gather_issue": float(res[3]), "cascade": float(res[4])},
                          "wall_ms_per_view": float(res[5]), "rank0_cascade_phases_ms_last_step": phases}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
