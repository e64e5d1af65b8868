"""Name registry mirroring the reference's plugin seam (modules/__init__.py:12-26): whatever class is visible as
``modules.<Name>`` is what a YAML ``block_type`` / ``fusion_function`` / ``classifier`` selects."""
import sys

from .mixer import (FeedForward, FusionMixer, MixerBlock, MLPMixer, MLPMixerNoPatching, PNLPMixer,  # noqa: F401
                    get_default_precision, set_default_precision)
from .fusion import BiModalGatedUnit, ConcatFusion, MaxFusion, MeanFusion, SumFusion  # noqa: F401
from .classification import BasicClassifier, MultilayerClassifier, StandardClassifier  # noqa: F401
from .mlp import MLP  # noqa: F401


def get_block_by_name(**kwargs):
    return getattr(sys.modules[__name__], kwargs['block_type'])(**kwargs)


def get_fusion_by_name(**kwargs):
    return getattr(sys.modules[__name__], kwargs['fusion_function'])(**kwargs)


def get_classifier_by_name(**kwargs):
    return getattr(sys.modules[__name__], kwargs['classifier'])(**kwargs)
