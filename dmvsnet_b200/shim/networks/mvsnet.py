"""Drop-in for the reference's ``networks/mvsnet.py``: put ``dmvsnet_b200/shim`` ahead of the reference
checkout on ``sys.path`` and ``from networks.mvsnet import MVSNet`` (reference model.py:9) resolves here."""
from dmvsnet_b200.module import *  # noqa: F401,F403  (the reference does `from .module import *`, mvsnet.py:6)
from dmvsnet_b200.mvsnet import Align_Corners_Range, CostAgg, DepthNet, MVSNet  # noqa: F401
