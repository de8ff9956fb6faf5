"""pidnet_b200 -- B200-native (sm_100a) execution engine for PIDNet behind the reference's
`models/pidnet.py` surface.  See DESIGN.md / INTEGRATION.md."""
from .pidnet import PIDNet, get_pred_model, get_seg_model  # noqa: F401
from .criterion import BondaryLoss, FullModel, OhemCrossEntropy  # noqa: F401
from .optim import FusedSGD, adjust_learning_rate  # noqa: F401

__all__ = ['PIDNet', 'get_pred_model', 'get_seg_model', 'OhemCrossEntropy', 'BondaryLoss', 'FullModel', 'FusedSGD',
           'adjust_learning_rate']
