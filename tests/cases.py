"""Shared case definitions: the inputs of every golden fixture are regenerated from these seeds."""
import torch

from dmvsnet_b200 import synthetic as syn

CASES = {
    "cascade_lin": dict(H=64, W=96, views=3, ndepths=[16, 8, 8], ratios=[4, 2, 1], inverse=False, batch=1, mode="features"),
    "cascade_inv_b2": dict(H=32, W=64, views=4, ndepths=[8, 8, 8], ratios=[4, 2, 1], inverse=True, batch=2, mode="features"),
    "cfg1_full": dict(H=128, W=160, views=4, ndepths=[48], ratios=[4], inverse=False, batch=1, mode="images"),
}


def case_inputs(case, seed=0):
    H, W, N, B = case["H"], case["W"], case["views"], case["batch"]
    ns = len(case["ndepths"])
    proj = syn.make_proj_matrices(H, W, N, B, num_stages=ns)
    dv = syn.make_depth_values(B, 192, inverse=case["inverse"])
    if case["mode"] == "images":
        return dict(imgs=syn.make_images(H, W, N, B, seed=seed), proj=proj, depth_values=dv)
    return dict(features=syn.make_stage_features(H, W, N, B, seed=seed, num_stages=ns), proj=proj, depth_values=dv)


def reference_state_keys():
    """state_dict of our MVSNet mirror: identical keys/shapes to the reference's (checked in test_host_cpu)."""
    from dmvsnet_b200 import MVSNet
    return MVSNet([8, 8, 8], [4, 2, 1]).state_dict()


def case_state(case, seed=0):
    from dmvsnet_b200 import MVSNet
    net = MVSNet(case["ndepths"], case["ratios"], inverse_depth=case["inverse"])
    return syn.randomise_regnet_state(net.state_dict(), seed=seed)


def rt_from_proj(proj_stage):
    from dmvsnet_b200 import ops
    return ops.relative_projections(proj_stage)
