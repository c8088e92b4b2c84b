#!/usr/bin/env python
"""bench.py - views/sec of the DMVSNet cost-volume hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config dtu|bmvs|tnt|synth|tiny]

One step = one pass of the hot path (3-stage cascade: S1 -> W1 -> R1 -> E1 -> W1 -> R1 -> E2 per stage) over one
synthetic DTU-shaped view set (1600x1184, N=5, D=[48,32,8]; BASELINE.json configs[1]).

  value      views/s, whole job, per-view features already resident in HBM (scope H of SURVEY 8d), CUDA-event timed
  e2e        views/s through the public API MVSNet.infer_many(): per step pinned host images -> H2D -> FeatureNet
             (dmvs_conv2d_f32) -> hot path -> D2H of depth + confidence (scope F + copies; the copies of neighbouring steps
             overlap compute on a copy stream); e2e.single_request_ms = one blocking MVSNet.infer() per step
  roofline   the fused warp+corr kernel (W1): algorithmic bytes 4*h*w*(N*C + 3*D) per launch over its CUDA-event time,
             all six passes of a step pooled; per-pass numbers under "roofline_per_launch"
  cpu_baseline / --impl reference
             the oracle port of the reference's PyTorch-CPU path (oracle/dmvs_oracle.py; the reference itself is pure
             Python and cannot travel to the GPU box) on the host cores, on a bounded sample (a 1600x160 band of the
             same view set, all stages, incl. FeatureNet), scaled linearly in rows to a full view.

N > 1: launched by torchrun, one process per GPU, independent replicas (one view set each; the path has no
data-path collective - DESIGN.md "Multi-GPU"), barrier + max-over-ranks timing, scaling "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (H, W, views, ndepths)
    "tiny": (128, 160, 4, [48]),
    "dtu": (1184, 1600, 5, [48, 32, 8]),
    "bmvs": (576, 768, 7, [48, 32, 8]),
    "tnt": (1056, 1920, 11, [48, 32, 8]),
    "synth": (3072, 4096, 9, [64, 32, 16]),
}
RATIOS = {1: [4], 3: [4, 2, 1]}
FEATURE_C = (32, 16, 8)
CPU_BAND_ROWS = 160


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    return 6650.0, "fallback (B200_PROFILING.md)", {}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons with nvidia-smi while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_band_step(state, ndepths, ratios, views, width, rows, seed=0):
    import torch
    from dmvsnet_b200 import synthetic as syn
    from oracle import dmvs_oracle as O
    imgs = syn.make_images(rows, width, views, 1, seed=seed, natural=True)
    proj = syn.make_proj_matrices(rows, width, views, 1, num_stages=len(ndepths))
    dv = syn.make_depth_values(1, 192, inverse=True)

    def step():
        with torch.no_grad():
            out = O.mvsnet_forward(imgs, proj, dv, state, ndepths, ratios, inverse_depth=True)
        return float(out["depth"].mean())
    return step


def cpu_state(ndepths, ratios):
    from dmvsnet_b200 import MVSNet, synthetic as syn
    net = MVSNet(ndepths, ratios, inverse_depth=True)
    return syn.randomise_regnet_state(net.state_dict(), seed=0)


def run_reference_arm(args, cfg_name):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    H, W, views, ndepths = CONFIGS[cfg_name]
    ratios = RATIOS[len(ndepths)]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rows = min(CPU_BAND_ROWS, H)
    step = cpu_band_step(cpu_state(ndepths, ratios), ndepths, ratios, views, W, rows)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    scale = H / float(rows)
    value = 1.0 / (dt * scale)
    sample = "%dx%d band (%d of %d rows) of the same view set, all stages incl. FeatureNet, scaled x%.2f in rows" % (W, rows, rows, H, scale)
    line = {
        "impl": "reference", "metric": "views/sec", "value": value, "unit": "views/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * scale * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg_name), "scope": "MVSNet.forward incl. FeatureNet, PyTorch-CPU fp32", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "views/s", "cores": cores, "kind": "port", "sample": sample,
                         "cpu": cpu_model(), "torch_threads": torch.get_num_threads()},
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def cpu_model():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def workload_name(cfg_name):
    H, W, views, nd = CONFIGS[cfg_name]
    return "%s %dx%d N=%d D=%s, B=1, 3-stage cascade, inverse_depth" % (cfg_name.upper(), W, H, views, nd)


# ------------------------------------------------------------------------------------------ GPU arm
def w1_bytes(h, w, n_views, c, d):
    return 4 * h * w * (n_views * c + 3 * d)


def in_bounds_fraction(rt, hyp, step=4):
    """Fraction of plane-sweep samples whose bilinear footprint centre lies inside the source image (subsampled)."""
    import torch
    b, d, h, w = hyp.shape
    ys = torch.arange(0, h, step, device=hyp.device, dtype=torch.float32)
    xs = torch.arange(0, w, step, device=hyp.device, dtype=torch.float32)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    dep = hyp[:, :, ::step, ::step]
    fr = []
    for s in range(rt.shape[1]):
        m = rt[0, s]
        X = (m[0] * xx + m[1] * yy + m[2]) * dep + m[9]
        Y = (m[3] * xx + m[4] * yy + m[5]) * dep + m[10]
        Z = (m[6] * xx + m[7] * yy + m[8]) * dep + m[11]
        u, v = X / Z, Y / Z
        fr.append(((u >= 0) & (u <= w - 1) & (v >= 0) & (v <= h - 1)).float().mean())
    return float(torch.stack(fr).mean())


def run_gpu_arm(args, cfg_name):
    import torch
    import torch.distributed as dist
    from dmvsnet_b200 import MVSNet, _native, ops, synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (one process per GPU); see the module docstring" % args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator comes up; the contract is ONE JSON line there
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _native.load()

    H, W, views, ndepths = CONFIGS[cfg_name]
    ratios = RATIOS[len(ndepths)]
    net = MVSNet(ndepths, ratios, inverse_depth=True)
    state = syn.randomise_regnet_state(net.state_dict(), seed=0)
    net.load_state_dict(state)
    net = net.to(dev).eval()
    net.DepthNet.return_prob_volume = True  # reference default: the probability volumes are part of the output

    # every rank works on its own view set (independent replicas)
    imgs_host = syn.make_images(H, W, views, 1, seed=rank, natural=True).pin_memory()
    proj = syn.make_proj_matrices(H, W, views, 1, num_stages=len(ndepths))
    dv_host = syn.make_depth_values(1, 192, inverse=True)
    dv = dv_host.to(dev)
    torch.backends.cudnn.benchmark = True  # reference model.py:25

    with torch.no_grad():
        imgs_dev = imgs_host.to(dev)
        feats = net.extract_features(imgs_dev)
        del imgs_dev
    torch.cuda.synchronize()

    def hot_step():
        with torch.no_grad():
            return net.cascade(feats, proj, dv, (H, W))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also JIT-free: everything is precompiled), then the timed hot-path region
    for _ in range(max(args.warmup, 3)):
        out = hot_step()
    del out
    barrier()
    if args.profile_step:
        # exactly one hot-path step between cudaProfilerStart/Stop, for `ncu --profile-from-start off`
        torch.cuda.cudart().cudaProfilerStart()
        hot_step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return 0
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ops.PROFILE = []
    launches0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = hot_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _native.launch_count() - launches0
    prof, ops.PROFILE = ops.PROFILE, None
    clocks = sampler.stop()
    depth_mean = float(out["depth"].mean())

    # in-bounds fraction of the six W1 launches (one extra untimed step, recorded hypotheses)
    inb = {}
    ops.CAPTURE = []
    hot_step()
    for tag, rt_t, hyp_t in ops.CAPTURE:
        inb[tag] = in_bounds_fraction(rt_t, hyp_t)
    ops.CAPTURE = None
    del out

    # ---- e2e: host images in, host depth/confidence out, every step.  Two numbers: `e2e` = the streaming loop a test run
    # executes (Model.test visits one reference view after the other): MVSNet.infer_many, where item k+1's H2D and item
    # k-1's D2H run on a copy stream under item k's compute - every step still performs its own copies inside the timed
    # region; `e2e_single` = one blocking MVSNet.infer() call per step (latency of a lone request).
    for _ in range(2):
        host = net.infer(imgs_host, proj, dv_host)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0.record()
    for _ in range(e2e_steps):
        host = net.infer(imgs_host, proj, dv_host)
    t1.record()
    barrier()
    e2e_single_ms = t0.elapsed_time(t1)
    for host in net.infer_many([(imgs_host, proj, dv_host)] * 4):  # warm-up: fills the pipeline, so every pinned result buffer exists
        pass
    # two passes of e2e_steps items each, the faster one is reported (both are listed): a single host-side hiccup (pinned
    # allocator growth, a page-fault burst on a fresh box) otherwise lands on 10 steps
    e2e_passes = []
    for _ in range(2):
        barrier()
        wall0 = time.perf_counter()
        t0.record()
        n_out = 0
        for host in net.infer_many([(imgs_host, proj, dv_host)] * e2e_steps):
            n_out += 1
        t1.record()
        barrier()
        assert n_out == e2e_steps
        # t1 is recorded after the generator has waited for the last D2H
        e2e_passes.append((t0.elapsed_time(t1), (time.perf_counter() - wall0) * 1e3))
    e2e_ms, e2e_wall_ms = min(e2e_passes)
    h2d = imgs_host.numel() * 4 + dv_host.numel() * 4 + sum(v.numel() * 4 for v in proj.values())
    d2h = sum(v.numel() * 4 for v in host.values())

    # ---- device-resident full forward (scope F without the copies), for the breakdown
    imgs_dev = imgs_host.to(dev)
    with torch.no_grad():
        net(imgs_dev, proj, dv)
        barrier()
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(3):
            net(imgs_dev, proj, dv)
        f1.record()
        barrier()
    full_ms = f0.elapsed_time(f1) / 3

    # ---- reduce over ranks (max time)
    t = torch.tensor([ms, e2e_ms, full_ms, e2e_single_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, full_ms, e2e_single_ms = [float(x) for x in t]

    # ---- W1 roofline from the per-launch events recorded inside the timed region
    peak, peak_src, _ = peaks()
    per_launch = {}
    for (tag, ev0, ev1, nbytes) in prof:
        d = per_launch.setdefault(tag, [0.0, 0, nbytes])
        d[0] += ev0.elapsed_time(ev1); d[1] += 1
    n_steps = float(args.steps)
    w1 = {k: v for k, v in per_launch.items() if k.startswith("w1:")}
    w1_ms = sum(v[0] for v in w1.values()) / n_steps
    w1_bytes_total = sum(v[2] * v[1] for v in w1.values()) / n_steps
    roof_rows = []
    for i, (k, v) in enumerate(sorted(w1.items())):
        gbs = v[2] / (v[0] / v[1] * 1e-3) / 1e9
        roof_rows.append({"launch": k, "ms": v[0] / v[1], "alg_MB": v[2] / 1e6, "GBps": gbs, "frac": gbs / peak,
                          "in_bounds": inb.get(k)})
    achieved = w1_bytes_total / (w1_ms * 1e-3) / 1e9 if w1_ms > 0 else 0.0
    # What actually bounds a fused bilinear gather: every sample pulls 4 corners x C floats through the SM's
    # L1 / shared-memory data path (128 B per clock per SM), whatever HBM does.  Tags are "w1:C<c>_D<d>_<h>x<w>".
    gather_bytes = 0.0
    for k in w1:
        c_, d_, hw_ = k[3:].split("_")
        h_, w_ = hw_.split("x")
        gather_bytes += float(h_) * float(w_) * int(d_[1:]) * (views - 1) * 4 * int(c_[1:]) * 4
    sm_peak = 148 * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9  # GB/s
    traffic = None
    tp = os.path.join(ROOT, "profiles", "w1_traffic.json")
    if os.path.exists(tp) and cfg_name == "dtu":
        traffic = json.load(open(tp)).get("dram_bytes_per_step")
    groups = {}
    for k, v in per_launch.items():
        if not k.startswith("w1:"):
            groups[k.split(":")[0]] = groups.get(k.split(":")[0], 0.0) + v[0] / n_steps

    if rank == 0:
        line = {
            "metric": "views/sec", "value": world * args.steps / (ms * 1e-3), "unit": "views/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg_name),
                       "arithmetic": "W1 / heads / sampler fp32; regularisation nets on tcgen05 kind::f16 with hi/lo-split fp16 operands "
                                     "(hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) = fp32-class accuracy; FeatureNet fp32 direct convolutions (dmvs_conv2d_f32)",
                       "scope_value": "hot path (stage loop mvsnet.py:208-258), features resident in HBM",
                       "scope_e2e": "MVSNet.infer_many: per step pinned host imgs -> H2D -> FeatureNet (dmvs_conv2d_f32) -> hot path -> D2H depth+confidence",
                       "l2": "inputs (1.06 GB of features + >1 GB of activations per step) exceed the 126 MB L2; no flush needed",
                       "parallelism": "replicas x%d (one view set per GPU, no collective)" % world, "weights": "random (SURVEY App. D recipe)",
                       "prob_volume": "kept (reference default)"},
            "clocks": clocks,
            "e2e": {"value": world * e2e_steps / (e2e_ms * 1e-3), "unit": "views/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "api": "MVSNet.infer_many (streaming: copies of neighbouring steps overlap compute)",
                    "wall_ms_per_step": e2e_wall_ms / e2e_steps, "passes_ms_per_step": [p[0] / e2e_steps for p in e2e_passes], "single_request_ms": e2e_single_ms / e2e_steps, "single_request_views_per_s": world * e2e_steps / (e2e_single_ms * 1e-3)},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "W1 = warp_corr_staged_kernel (stage-1 planes) + warp_corr_nhwc_kernel (regressed hypotheses), 6 passes / 7 launches per step pooled", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": w1_bytes_total, "ms_per_step": w1_ms,
                         "on_chip": {"what": "bytes the gather moves through the SMs' L1/shared-memory data path (samples x 4 corners x C x 4 B) "
                                             "against 148 SMs x 128 B/clk at the sampled SM clock: the floor of any fp32 gather formulation",
                                     "bytes_per_step": gather_bytes, "peak_GBps": sm_peak, "floor_ms": gather_bytes / sm_peak / 1e6,
                                     "frac": (gather_bytes / sm_peak / 1e6) / w1_ms if w1_ms > 0 else None}},
            "roofline_per_launch": roof_rows,
            "breakdown_ms_per_step": dict(groups, w1=w1_ms),
            "full_forward_device_ms": full_ms,
            "check": {"depth_mean": depth_mean},
        }
        if world == 1 and not args.no_cpu:
            import torch as _t
            cores = os.cpu_count() or 1
            _t.set_num_threads(cores)
            rows = min(CPU_BAND_ROWS, H)
            step = cpu_band_step(state, ndepths, ratios, views, W, rows)
            step()
            t0c = time.perf_counter(); step(); dtc = time.perf_counter() - t0c
            scale = H / float(rows)
            line["cpu_baseline"] = {"value": 1.0 / (dtc * scale), "unit": "views/s", "cores": cores, "kind": "port", "cpu": cpu_model(),
                                    "sample": "%dx%d band (%d of %d rows), all stages incl. FeatureNet, 1 warm-up + 1 timed pass, scaled x%.2f"
                                              % (W, rows, rows, H, scale), "seconds_for_sample": dtc}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="dtu", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--profile-step", action="store_true", help="run one hot-path step inside cudaProfilerStart/Stop and exit (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args, args.config)
    return run_gpu_arm(args, args.config)


if __name__ == "__main__":
    sys.exit(main())
