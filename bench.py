#!/usr/bin/env python
"""bench.py - views/sec of DMVSNet's cost-volume path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config dtu|bmvs|tnt|synth|tiny]

One step = one reference view of BASELINE.json configs[1] (DTU 1600x1184, N=5, D=[48,32,8]): MVSNet.forward = FeatureNet on
the N images, then the 3-stage cascade (S1 -> W1 -> R1 -> E1 -> W1 -> R1 -> E2 per stage).

  value      views/s of the full forward with the images already resident in HBM (scope F of SURVEY 8d, no copies), CUDA-event
             timed, whole job over all ranks.  The reference arm (--impl reference) reports the same scope on the host CPU.
  e2e        views/s through the public API MVSNet.infer_many(): per step pinned host images -> H2D -> FeatureNet -> cascade ->
             D2H of depth + confidence (copies of neighbouring steps overlap compute on a copy stream; every step still pays its
             own); e2e.single_request_ms = one blocking MVSNet.infer() per step
  hot_path   the stage loop alone (scope H) on feature maps resident in HBM, on the CONDITIONED workload: photo-consistent
             feature maps of a tilted plane (synthetic.make_scene_features) - what a trained FeatureNet hands the cascade -
             so that the regressed depth maps, hence W1's gather pattern, are the piecewise-smooth ones of a trained network
  roofline   the fused warp+corr kernel W1 inside the hot_path region: algorithmic bytes 4*h*w*(N*C + 3*D) per launch over
             its CUDA-event time, six launches per step pooled, against the measured HBM peak; per launch under
             "roofline_per_launch".  "roofline_full_forward": the same kernel inside the `value` region, where the feature maps
             come from FeatureNet itself (random weights; output heads calibrated on the synthetic images unless --featnet-heads random)
  cpu_baseline / --impl reference
             the oracle port of the reference's PyTorch-CPU path (oracle/dmvs_oracle.py; the reference itself is pure Python and
             does not exist on the GPU box) on all host cores, on the SAME full view (no band, no scaling)
  gpu_library_baseline
             the same restatement with its tensors on the B200 (the reference's PyTorch-CUDA / cuDNN path), TF32 off and on
  check      final depth of the benchmarked view against the oracle's (computed once, untimed)

Weights: FeatureNet random with its three bias-free output heads calibrated to zero-mean, unit-variance descriptors
on the synthetic images (synthetic.calibrate_feature_heads: without it the correlation is dominated by the common mode of the post-ReLU
activations, the cost volume has no ridge and the cascade regresses noise), regularisation nets synthetic.ridge_regnet_state (follow the cost ridge like trained
ones; all layers dense).  N > 1: torchrun, one process per GPU, independent replicas (one view set each; the path has no
data-path collective - DESIGN.md "Multi-GPU"), barrier + max-over-ranks timing, scaling "weak"; rank 0 additionally times the
single-view sharded mode when --sharded is given.
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
# algorithmic flops of the regularisation nets per view (2 * MAC): main + refine nets of all stages
R1_GFLOP = {"dtu": 708.7, "bmvs": 165.5, "tnt": 758.5, "synth": 6468.6}


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


# ------------------------------------------------------------------------------------------ workload (both arms)
FEATNET_HEADS = "calibrated"  # --featnet-heads: "calibrated" (synthetic.calibrate_feature_heads) | "random"


def make_workload(cfg_name, seed, want_features=True):
    """The inputs of one step: images of the rendered scene, cameras, depth range, the network state - identical for both arms -
    and (GPU arm) the conditioned feature maps of the hot_path region."""
    import torch
    from dmvsnet_b200 import MVSNet, synthetic as syn
    H, W, views, ndepths = CONFIGS[cfg_name]
    ratios = RATIOS[len(ndepths)]
    proj = syn.make_proj_matrices(H, W, views, 1, num_stages=len(ndepths))
    dv = syn.make_depth_values(1, 192, inverse=True)
    imgs = syn.make_scene_images(H, W, views, proj["stage%d" % len(ndepths)], seed=seed)
    # photographs are 8-bit: the float images every arm sees are u8 / 255 (what the reference's loaders compute, general_eval.py:161)
    imgs_u8 = (imgs * 255.0).round().clamp_(0, 255).to(torch.uint8)
    imgs = imgs_u8.to(torch.float32) / 255.0
    net = MVSNet(ndepths, ratios, inverse_depth=True)
    state = syn.ridge_regnet_state(net.state_dict(), seed=0)
    if FEATNET_HEADS == "calibrated":
        state = syn.calibrate_feature_heads(state, imgs)
    feats = syn.make_scene_features(H, W, views, proj, seed=seed, num_stages=len(ndepths)) if want_features else None
    return dict(H=H, W=W, views=views, ndepths=ndepths, ratios=ratios, proj=proj, dv=dv, imgs=imgs, imgs_u8=imgs_u8, net=net, state=state,
                feats=feats)


def workload_name(cfg_name):
    H, W, views, nd = CONFIGS[cfg_name]
    return "%s %dx%d N=%d D=%s, B=1, 3-stage cascade, inverse_depth" % (cfg_name.upper(), W, H, views, nd)


def cpu_model():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def oracle_forward(wl, device=None):
    """MVSNet.forward of the oracle port on ``device`` (None: host CPU).  Returns a closure that runs one full view."""
    import torch
    from oracle import dmvs_oracle as O
    if device is None:
        imgs, proj, dv, state = wl["imgs"], wl["proj"], wl["dv"], wl["state"]
    else:
        imgs, dv = wl["imgs"].to(device), wl["dv"].to(device)
        proj = {k: v.to(device) for k, v in wl["proj"].items()}
        state = {k: v.to(device) for k, v in wl["state"].items()}

    def step():
        with torch.no_grad():
            return O.mvsnet_forward(imgs, proj, dv, state, wl["ndepths"], wl["ratios"], inverse_depth=True)
    return step


def depth_errors(got, want):
    """max / p99.9 / mean relative error and the fraction of pixels beyond the 1e-3 contract."""
    e = ((got.detach().cpu().double() - want.detach().cpu().double()).abs() / want.detach().cpu().double().abs().clamp_min(1.0)).flatten()
    k = max(1, int(0.999 * e.numel()))
    return {"max": float(e.max()), "p99_9": float(e.kthvalue(k)[0]), "mean": float(e.mean()), "frac_beyond_1e-3": float((e > 1e-3).double().mean())}


def run_reference_arm(args, cfg_name):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads, the same full view."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = make_workload(cfg_name, seed=0, want_features=False)
    step = oracle_forward(wl)
    # a full DTU view is ~25 s on 16 cores: the step count is capped so that the run ends within a few minutes
    warm, timed = min(args.warmup, 1), max(1, min(args.steps, 2))
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(timed):
        out = step()
    dt = (time.perf_counter() - t0) / timed
    value = 1.0 / dt
    sample = "the full %dx%d view (no band, no scaling), %d warm-up + %d timed passes (of --steps %d --warmup %d asked: one pass is ~%.0f s)" % (
        wl["W"], wl["H"], warm, timed, args.steps, args.warmup, dt)
    line = {
        "impl": "reference", "metric": "views/sec", "value": value, "unit": "views/s", "n_gpus": args.gpus, "steps": timed,
        "warmup": warm, "steps_requested": args.steps, "warmup_requested": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cfg_name), "scope_value": "MVSNet.forward incl. FeatureNet (scope F), PyTorch-CPU fp32", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "views/s", "cores": cores, "kind": "port", "sample": sample,
                         "cpu": cpu_model(), "torch_threads": torch.get_num_threads()},
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "check": {"depth_mean": float(out["depth"].mean())},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ GPU arm
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


def w1_roofline(prof, n_steps, views, peak, inb=None):
    """Pool the W1 launches recorded by ops.PROFILE: (summary dict, per-launch rows, other groups' ms per step)."""
    per_launch = {}
    for (tag, ev0, ev1, nbytes) in prof:
        d = per_launch.setdefault(tag, [0.0, 0, nbytes])
        d[0] += ev0.elapsed_time(ev1); d[1] += 1
    w1 = {k: v for k, v in per_launch.items() if k.startswith("w1:")}
    w1_ms = sum(v[0] for v in w1.values()) / n_steps
    w1_bytes_total = sum(v[2] * v[1] for v in w1.values()) / n_steps
    rows = []
    for k, v in sorted(w1.items()):
        gbs = v[2] / (v[0] / v[1] * 1e-3) / 1e9
        rows.append({"launch": k, "ms": v[0] / v[1], "alg_MB": v[2] / 1e6, "GBps": gbs, "frac": gbs / peak, "in_bounds": (inb or {}).get(k)})
    achieved = w1_bytes_total / (w1_ms * 1e-3) / 1e9 if w1_ms > 0 else 0.0
    # what the gather moves through the SMs' L1 / shared-memory data pipe: samples x 4 corners x C x 2 B (fp16 sources)
    gather_bytes = 0.0
    for k in w1:
        c_, d_, hw_ = k[3:].split("_")
        h_, w_ = hw_.split("x")
        gather_bytes += float(h_) * float(w_) * int(d_[1:]) * (views - 1) * 4 * int(c_[1:]) * 2
    groups = {}
    for k, v in per_launch.items():
        if not k.startswith("w1:"):
            groups[k.split(":")[0]] = groups.get(k.split(":")[0], 0.0) + v[0] / n_steps
    return {"achieved": achieved, "frac": achieved / peak, "algorithmic_bytes_per_step": w1_bytes_total, "ms_per_step": w1_ms,
            "gather_bytes_per_step": gather_bytes}, rows, dict(groups, w1=w1_ms)


def sharded_leg(args, world, rank, dev, barrier):
    """Single-view mode (BASELINE config 4): ONE view set on all `world` GPUs - FeatureNet by view + all-gather of the
    fp16 source maps, the stage loop by row bands with a 32-row halo (dmvsnet_b200/parallel.py) - timed beside the same view on
    one GPU (every rank runs the unsharded forward on the same inputs; max over ranks of both)."""
    import torch
    import torch.distributed as dist
    from dmvsnet_b200 import parallel
    cfg = args.sharded_config
    wl = make_workload(cfg, seed=0, want_features=False)
    net = wl["net"]
    net.load_state_dict(wl["state"])
    net = net.to(dev).eval()
    net.DepthNet.return_prob_volume = False  # the sharded mode leaves the volumes on the ranks that own the rows
    net.w1_precision = "fp16"
    imgs = wl["imgs"].to(dev)
    proj, dv = wl["proj"], wl["dv"].to(dev)

    def single():
        with torch.no_grad():
            return net(imgs, proj, dv)

    def sharded():
        with torch.no_grad():
            return parallel.infer_view_sharded(net, imgs, proj, dv)

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / steps, out
    for _ in range(3):
        single(); sharded()
    steps = max(3, min(args.steps, 10))
    t1, ref = timed(single, steps)
    tn, out = timed(sharded, steps)
    same = all(torch.equal(out[k], ref[k]) for k in ("depth", "photometric_confidence")) and \
        all(torch.equal(out["stage%d" % s]["depth"], ref["stage%d" % s]["depth"]) for s in (1, 2, 3))
    worst = max(float((out["stage%d" % s]["depth"] - ref["stage%d" % s]["depth"]).abs().max() / ref["stage%d" % s]["depth"].abs().max()) for s in (1, 2, 3))
    # the exchanged bytes of one view: fp16 source maps (all-gather) + the per-stage maps (all-reduce of zero-padded maps)
    t = torch.tensor([t1, tn, 0.0 if same else 1.0, worst], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    H, W, views = wl["H"], wl["W"], wl["views"]
    gather_bytes = sum(2 * 2 * c * (H >> (2 - s)) * (W >> (2 - s)) for s, c in enumerate(FEATURE_C)) * (views - 1)
    reduce_bytes = sum(4 * 15 * (H >> (2 - s)) * (W >> (2 - s)) for s in range(3))
    return {"workload": workload_name(cfg), "ranks": world, "ms_per_view_1gpu": float(t[0]), "ms_per_view": float(t[1]),
            "speedup_vs_1gpu": float(t[0] / t[1]), "bit_identical": bool(t[2] == 0.0), "depth_max_rel_diff": float(t[3]),
            "scope": "MVSNet.forward incl. FeatureNet, images resident on every GPU, depth + confidence re-assembled on every GPU",
            "exchange": {"all_gather_fp16_source_maps_bytes": gather_bytes, "all_gather_per_pixel_maps_bytes": reduce_bytes,
                         "collectives_per_view": 6 + 2 * 3},
            "partition": "FeatureNet by view (+ the reference view on every rank); stage loop by row bands, 32-row halo, edges at multiples of 8 rows"}


def run_gpu_arm(args, cfg_name):
    import torch
    import torch.distributed as dist
    from dmvsnet_b200 import _native, ops

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

    # every rank works on its own view set (independent replicas)
    wl = make_workload(cfg_name, seed=rank)
    H, W, views, ndepths = wl["H"], wl["W"], wl["views"], wl["ndepths"]
    net = wl["net"]
    net.load_state_dict(wl["state"])
    net = net.to(dev).eval()
    net.DepthNet.return_prob_volume = True  # reference default: the probability volumes are part of the output
    net.w1_precision = args.w1
    imgs_host = wl["imgs"].pin_memory()
    proj, dv_host = wl["proj"], wl["dv"]
    dv = dv_host.to(dev)
    imgs_dev = imgs_host.to(dev)
    with torch.no_grad():
        cond = [{k: v.to(dev) for k, v in f.items()} for f in wl["feats"]]
        if net.w1_precision == "fp16":
            net.add_half_features(cond)  # what FeatureNet's epilogue emits for the source views
    torch.cuda.synchronize()

    def full_step():
        with torch.no_grad():
            return net(imgs_dev, proj, dv)

    def hot_step():
        with torch.no_grad():
            return net.cascade(cond, proj, dv, (H, W))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=None, settle=2):
        # profile: None = no per-op events; "w1:" = CUDA events around the W1 launches only (12 per step: the roofline is measured
        # inside the timed region); "" = around every op (host time: used in a separate short pass for the breakdown).
        # `settle` untimed iterations of the SAME loop body first: the caching allocator reaches the block layout of this loop
        # (a cudaMalloc inside the timed region is a device-wide stall of milliseconds), and no result outlives its iteration.
        out = None
        for _ in range(settle):
            out = None
            out = fn()
        if profile is not None:
            ops.PROFILE, ops.PROFILE_ONLY = [], (profile or None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = _native.launch_count()
        e0.record()
        for _ in range(steps):
            out = None
            out = fn()
        e1.record()
        barrier()
        prof, ops.PROFILE, ops.PROFILE_ONLY = ops.PROFILE, None, None
        return e0.elapsed_time(e1), out, prof, _native.launch_count() - l0

    warm = max(args.warmup, 3)
    for _ in range(warm):
        full_step()
        hot_step()
    barrier()
    if args.profile_step:
        # exactly one step between cudaProfilerStart/Stop, for `ncu --profile-from-start off`
        torch.cuda.cudart().cudaProfilerStart()
        (full_step if args.profile_step == "full" else hot_step)()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return 0
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    # ---- value: the full forward, images resident in HBM; CUDA events around the W1 launches ride along
    ms, out, prof_full, launches = timed(full_step, args.steps, profile="w1:")
    depth_full = out["depth"].clone()
    # ---- hot_path: the stage loop on the conditioned feature maps
    hot_ms, out, prof_hot, hot_launches = timed(hot_step, args.steps, profile="w1:")
    depth_hot = out["depth"].clone()
    clocks = sampler.stop()
    del out
    # per-group breakdown: a separate short pass with events around every op (not part of any reported throughput)
    _, _, prof_full_all, _ = timed(full_step, 3, profile="")
    _, _, prof_hot_all, _ = timed(hot_step, 3, profile="")

    # in-bounds fraction of the six W1 launches (one extra untimed step each, recorded hypotheses)
    inb_hot, inb_full = {}, {}
    for fn, dst in ((hot_step, inb_hot), (full_step, inb_full)):
        ops.CAPTURE = []
        fn()
        for tag, rt_t, hyp_t in ops.CAPTURE:
            dst[tag] = in_bounds_fraction(rt_t, hyp_t)
        ops.CAPTURE = None

    # ---- e2e: host images in, host depth/confidence out, every step.  Headline: the 8-bit photographs cross PCIe as they are and
    # are scaled by 1/255 on the device (MVSNet.infer's uint8 entry); "float32_images": the reference's own hand-over (fp32 images
    # made by the loader on the host), 4x the H2D bytes.
    imgs_u8_host = wl["imgs_u8"].pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def e2e_of(host_imgs):
        for _ in range(2):
            host = net.infer(host_imgs, proj, dv_host)
        single_ms, _, _, _ = timed(lambda: net.infer(host_imgs, proj, dv_host), e2e_steps)
        for host in net.infer_many([(host_imgs, proj, dv_host)] * 4):  # warm-up: fills the pipeline, every pinned result buffer exists
            pass
        # two passes of e2e_steps items each, the faster one is reported (both are listed): a single host-side hiccup (pinned
        # allocator growth, a page-fault burst on a fresh box) otherwise lands on 10 steps
        passes = []
        for _ in range(2):
            barrier()
            wall0 = time.perf_counter()
            t0.record()
            n_out = 0
            for host in net.infer_many([(host_imgs, proj, dv_host)] * e2e_steps):
                n_out += 1
            t1.record()
            barrier()
            assert n_out == e2e_steps
            passes.append((t0.elapsed_time(t1), (time.perf_counter() - wall0) * 1e3))
        h2d = host_imgs.numel() * host_imgs.element_size() + dv_host.numel() * 4 + sum(v.numel() * 4 for v in proj.values())
        d2h = sum(v.numel() * 4 for v in host.values())
        return min(passes) + (single_ms, passes, h2d, d2h, host)
    e2e_ms, e2e_wall_ms, e2e_single_ms, e2e_passes, h2d, d2h, host_u8 = e2e_of(imgs_u8_host)
    f32_ms, f32_wall_ms, f32_single_ms, f32_passes, f32_h2d, _, host_f32 = e2e_of(imgs_host)
    same_result = bool(torch.equal(host_u8["depth"], host_f32["depth"]))

    # ---- reduce over ranks (max time)
    t = torch.tensor([ms, hot_ms, e2e_ms, e2e_single_ms, f32_ms, f32_single_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, hot_ms, e2e_ms, e2e_single_ms, f32_ms, f32_single_ms = [float(x) for x in t]

    sharded = None
    if world > 1 and not args.no_sharded:
        del cond
        torch.cuda.empty_cache()
        sharded = sharded_leg(args, world, rank, dev, barrier)

    peak, peak_src, pk = peaks()
    n_steps = float(args.steps)
    roof_hot, rows_hot, _ = w1_roofline(prof_hot, n_steps, views, peak, inb_hot)
    roof_full, rows_full, _ = w1_roofline(prof_full, n_steps, views, peak, inb_full)
    _, _, groups_hot = w1_roofline(prof_hot_all, 3.0, views, peak)
    _, _, groups_full = w1_roofline(prof_full_all, 3.0, views, peak)
    sm_peak = 148 * 128 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e9  # GB/s through the SMs' shared-memory data pipe
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r2n_w1_traffic.json")
    if os.path.exists(tp) and cfg_name == "dtu":
        traffic = json.load(open(tp)).get("dram_bytes_per_step")

    if rank == 0:
        kernel = {"fp16": "W1 = warp_corr_h16_kernel (fp16 source maps staged by TMA; reference view, weights, sums fp32), 6 launches per step pooled",
                  "fp32": "W1 = warp_corr_staged_kernel (stage-1 planes) + warp_corr_nhwc_kernel (regressed hypotheses), 6 passes / 7 launches per step pooled"}[net.w1_precision]
        line = {
            "metric": "views/sec", "value": world * args.steps / (ms * 1e-3), "unit": "views/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg_name),
                       "arithmetic": "fp32 throughout, except: W1's SOURCE feature maps are stored as fp16 (w1_precision=%s; products and sums fp32); "
                                     "regularisation nets and FeatureNet's 3x3 layers on tcgen05 kind::f16 with hi/lo-split fp16 operands "
                                     "(hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) = fp32-class accuracy" % net.w1_precision,
                       "scope_value": "MVSNet.forward incl. FeatureNet (scope F), images resident in HBM",
                       "scope_e2e": "MVSNet.infer_many: per step pinned host imgs (uint8) -> H2D -> /255 -> FeatureNet -> cascade -> D2H depth+confidence",
                       "scope_hot_path": "stage loop mvsnet.py:208-258 (scope H) on photo-consistent feature maps resident in HBM",
                       "l2": "inputs (113 MB of images, 1.06 GB of features, >1 GB of activations per step) exceed the 126 MB L2; no flush needed",
                       "parallelism": "replicas x%d (one view set per GPU, no collective)" % world,
                       "weights": "FeatureNet random (SURVEY App. D)%s; regularisation nets follow the cost ridge like trained ones (synthetic.ridge_regnet_state)"
                                  % (", output heads calibrated to zero-mean unit-variance descriptors on the synthetic images (synthetic.calibrate_feature_heads)"
                                     if FEATNET_HEADS == "calibrated" else ""),
                       "images": "renderings of a textured tilted plane seen by the rig's cameras (synthetic.make_scene_images)",
                       "prob_volume": "kept (reference default)"},
            "clocks": clocks,
            "e2e": {"value": world * e2e_steps / (e2e_ms * 1e-3), "unit": "views/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "api": "MVSNet.infer_many (streaming: copies of neighbouring steps overlap compute)",
                    "host_images": "uint8 (8-bit photographs; x / 255 in fp32 on the device, the same values as the reference's host-side conversion)",
                    "wall_ms_per_step": e2e_wall_ms / e2e_steps, "passes_ms_per_step": [p[0] / e2e_steps for p in e2e_passes],
                    "single_request_ms": e2e_single_ms / e2e_steps, "single_request_views_per_s": world * e2e_steps / (e2e_single_ms * 1e-3),
                    "float32_images": {"value": world * e2e_steps / (f32_ms * 1e-3), "ms_per_step": f32_ms / e2e_steps, "h2d_bytes_per_step": f32_h2d,
                                       "single_request_ms": f32_single_ms / e2e_steps, "same_depth_as_uint8_entry": same_result}},
            "gpu_launches": int(launches),
            "hot_path": {"value": world * args.steps / (hot_ms * 1e-3), "unit": "views/s", "ms_per_step": hot_ms / args.steps,
                         "gpu_launches": int(hot_launches), "breakdown_ms_per_step": groups_hot},
            "roofline": dict({"kernel": kernel, "bound": "hbm", "peak": peak, "unit": "GB/s", "traffic": traffic, "peak_source": peak_src,
                              "workload": "hot_path region (conditioned feature maps: the regressed depth is piecewise smooth, as a trained network's)",
                              "on_chip": {"what": "bytes the gather moves through the SMs' shared-memory data pipe (samples x 4 corners x C x 2 B) against 148 SMs x "
                                                  "128 B/clk at the sampled SM clock: the floor of this gather formulation; the kernel is instruction-issue bound above it",
                                          "peak_GBps": sm_peak, "floor_ms": roof_hot["gather_bytes_per_step"] / sm_peak / 1e6,
                                          "frac": (roof_hot["gather_bytes_per_step"] / sm_peak / 1e6) / roof_hot["ms_per_step"] if roof_hot["ms_per_step"] > 0 else None}},
                             **{k: v for k, v in roof_hot.items() if k != "gather_bytes_per_step"}),
            "roofline_per_launch": rows_hot,
            "roofline_regnet": None if cfg_name not in R1_GFLOP else {
                "kernel": "R1 = conv_tc2_kernel / conv_kf_kernel (tcgen05 implicit GEMM, fp16 hi/lo split operands: 3 products per multiply-add), all regularisation "
                          "nets of a step (77 % of the hot path's kernel time, profiles/r2m_launches_hot.csv)",
                "bound": "tensor", "unit": "TFLOP/s", "algorithmic_gflop_per_step": R1_GFLOP[cfg_name],
                "ms_per_step": groups_hot.get("regnet", 0.0) + groups_hot.get("regnet_refine", 0.0),
                "achieved": R1_GFLOP[cfg_name] / max(groups_hot.get("regnet", 0.0) + groups_hot.get("regnet_refine", 0.0), 1e-9),
                "peak": float(pk.get("bf16_tflops_sustained", 1400.0)),
                "frac": R1_GFLOP[cfg_name] / max(groups_hot.get("regnet", 0.0) + groups_hot.get("regnet_refine", 0.0), 1e-9) / float(pk.get("bf16_tflops_sustained", 1400.0)),
                "note": "algorithmic flops (one product per multiply-add) over the event-timed net launches against the sustained bf16 peak; the split executes 3x "
                        "these flops, and at base_channels = 8 the layers are bound by the shared-memory operand fetch per MMA (A 4 KB + B, ~68 B/clk: N = 16..96 "
                        "leaves the pipe 16-51 % busy) and by the per-plane hand-offs of the full-resolution layers, not by the tensor pipe (DESIGN.md section 4)"},
            "roofline_full_forward": dict({"workload": "value region: feature maps computed by FeatureNet from the rendered images (heads: %s)" % FEATNET_HEADS},
                                          **{k: v for k, v in roof_full.items() if k != "gather_bytes_per_step"}, per_launch=rows_full),
            "breakdown_ms_per_step": groups_full,
            "check": {"depth_mean": float(depth_full.mean()), "depth_mean_hot_path": float(depth_hot.mean())},
        }
        if sharded is not None:
            line["sharded"] = sharded
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            from oracle import dmvs_oracle as O
            # ---- cpu_baseline: the oracle port on the host cores, the same full view, one timed pass (the thread pool is warmed on
            # a small problem first).  Its depth map doubles as the checker of the benchmarked view.
            small = make_workload("tiny", seed=0, want_features=False)
            oracle_forward(small)()
            step = oracle_forward(wl)
            t0c = time.perf_counter()
            ref_out = step()
            dtc = time.perf_counter() - t0c
            line["cpu_baseline"] = {"value": 1.0 / dtc, "unit": "views/s", "cores": cores, "kind": "port", "cpu": cpu_model(),
                                    "sample": "the full %dx%d view, all stages incl. FeatureNet (scope F), 1 timed pass after warming the thread pool on a "
                                              "128x160 problem" % (W, H), "seconds_for_sample": dtc}
            line["check"]["depth_rel_err_vs_oracle"] = dict(depth_errors(depth_full, ref_out["depth"]),
                                                             what="final depth of the benchmarked full forward vs the oracle's on the same inputs")
            for sname in ("stage1", "stage2"):
                line["check"]["depth_rel_err_vs_oracle_" + sname] = depth_errors(full_step()[sname]["depth"], ref_out[sname]["depth"])
            del ref_out
            with torch.no_grad():
                ref_hot = O.cascade_forward(wl["feats"], proj, dv_host, wl["state"], ndepths, wl["ratios"], True, (H, W))
            line["check"]["depth_rel_err_vs_oracle_hot_path"] = dict(depth_errors(depth_hot, ref_hot["depth"]),
                                                                      what="final depth of the hot_path region vs the oracle's cascade on the same feature maps")
            del ref_hot
            # ---- gpu_library_baseline: the same restatement on the B200 through ATen / cuDNN (reference model.py:25)
            lib = {}
            torch.backends.cudnn.benchmark = True
            for name, tf32 in (("tf32_off", False), ("tf32_on", True)):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                try:
                    gstep = oracle_forward(wl, dev)
                    for _ in range(2):
                        gout = gstep()
                    gms, gout, _, _ = timed(gstep, 3)
                    lib[name] = {"value": 3 / (gms * 1e-3), "unit": "views/s", "ms_per_step": gms / 3,
                                 "depth_rel_err_vs_product": depth_errors(gout["depth"], depth_full)}
                    del gout
                except Exception as e:  # e.g. out of memory on the largest configuration
                    lib[name] = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
                torch.cuda.empty_cache()
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            lib["what"] = "oracle/dmvs_oracle.py (the reference's forward, restated) with its tensors on the B200: F.grid_sample, cuDNN conv3d / conv2d; images resident in HBM (scope F)"
            line["gpu_library_baseline"] = lib
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
    ap.add_argument("--w1", default="fp16", choices=["fp16", "fp32"], help="W1 source-map precision (MVSNet.w1_precision)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the single-view sharded leg")
    ap.add_argument("--sharded-config", default="tnt", choices=sorted(CONFIGS), help="N > 1: the view set of the single-view sharded leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / check / gpu_library_baseline legs")
    ap.add_argument("--featnet-heads", default="calibrated", choices=["calibrated", "random"],
                    help="FeatureNet's bias-free output heads: calibrated on the synthetic images (zero-mean, unit-variance descriptors, "
                         "synthetic.calibrate_feature_heads) or plain random")
    ap.add_argument("--profile-step", nargs="?", const="hot", default=None, choices=["hot", "full"],
                    help="run one step (hot path or full forward) inside cudaProfilerStart/Stop and exit (for ncu)")
    args = ap.parse_args()
    global FEATNET_HEADS
    FEATNET_HEADS = args.featnet_heads
    if args.impl == "reference":
        return run_reference_arm(args, args.config)
    return run_gpu_arm(args, args.config)


if __name__ == "__main__":
    sys.exit(main())
