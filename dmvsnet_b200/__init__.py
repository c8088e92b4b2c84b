"""dmvsnet_b200: B200-native (sm_100a) cost-volume hot path behind DMVSNet's ``networks.mvsnet`` API.

    from dmvsnet_b200 import MVSNet          # same constructor / forward / state_dict as the reference

The arithmetic lives in ``libdmvs_b200.so`` (hand-written CUDA, C ABI in ``include/dmvs_b200.h``);
build it with ``python -m dmvsnet_b200.build``.  See DESIGN.md / INTEGRATION.md.
"""
from .mvsnet import MVSNet, CostAgg, DepthNet  # noqa: F401
from .module import (CostRegNet, CostRegNet_refine, FeatureNet, get_depth_range_samples, homo_warping,  # noqa: F401
                     depth_regression)

__version__ = "0.1.0"
