"""aes_lac_2018_b200 -- B200-native CTC loss-and-gradient engine.

The one hot path of igormq/aes-lac-2018 (its `warpctc_pytorch.CTCLoss` call), rebuilt as hand-written
sm_100a CUDA behind the same Python and C interfaces.  See DESIGN.md.
"""
from .ctc_loss import CTCLoss, _CTC, ctc_loss_raw, ctc_loss_host, sanitize_loss, release_workspaces  # noqa: F401
from .decoder import GreedyDecoder, OrderedAlphabet, greedy_decode_raw, edit_distance_raw  # noqa: F401
from .metrics import WER, CER, EditDistance  # noqa: F401
from .head import SequenceWiseClassifier  # noqa: F401

__all__ = ["CTCLoss", "ctc_loss_raw", "ctc_loss_host", "sanitize_loss", "release_workspaces", "GreedyDecoder",
           "OrderedAlphabet", "greedy_decode_raw",
           "edit_distance_raw", "WER", "CER", "EditDistance", "SequenceWiseClassifier"]
__version__ = "0.1.0"
