"""Import alias so that the reference's own lines work unchanged:

    from warpctc_pytorch import CTCLoss as warp_CTCLoss     # /root/reference/train.py:12
    from warpctc_pytorch import CTCLoss as warp_CTCLoss     # /root/reference/codes/metrics.py:3

Put the repository root on PYTHONPATH (or pip-install it) and the reference trains against the
B200-native engine.  INTEGRATION.md shows the alternatives.
"""
from aes_lac_2018_b200 import CTCLoss, _CTC  # noqa: F401

__all__ = ["CTCLoss"]
