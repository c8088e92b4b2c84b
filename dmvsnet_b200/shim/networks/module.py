"""Drop-in for the reference's ``networks/module.py`` (see ``mvsnet.py`` next to this file)."""
from dmvsnet_b200.module import *  # noqa: F401,F403
from dmvsnet_b200.module import CostRegNet_part, CostRegNet_part_refine  # noqa: F401
