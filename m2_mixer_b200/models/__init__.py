"""``models.get_model(type)`` as in the reference (models/__init__.py:14-15)."""
import sys

from .avmnist import AVMnistMixerMultiLoss  # noqa: F401
from .mimic import MimicMixerMultiLoss  # noqa: F401
from .mmimdb import MMIMDBMixerMultiLoss  # noqa: F401


def get_model(model_type: str):
    return getattr(sys.modules[__name__], model_type)
