"""MimicMixerMultiLoss (reference models/mimic.py:24-142): static-tabular MLP encoder + time-series Mixer, batch is
the tuple (static [B,5], time [B,24,12], label [B])."""
from __future__ import annotations

import torch

from .. import functional as F
from .. import modules
from ..config import wrap
from .base import TrainTestModule


class MimicMixerMultiLoss(TrainTestModule):
    def __init__(self, model_cfg, optimizer_cfg=None, **kwargs):
        model_cfg = wrap(model_cfg)
        self.num_classes = model_cfg.modalities.classification.get('num_classes', 3)
        super().__init__(optimizer_cfg, **kwargs)
        self.modalities_freezed = False
        m = model_cfg.modalities
        dropout = model_cfg.get('dropout', 0.0)
        self.time_mixer = modules.get_block_by_name(**m.time, dropout=dropout)
        self.static_extractor = modules.get_block_by_name(**m.static, dropout=dropout)
        self.fusion_function = modules.get_fusion_by_name(**m.multimodal)
        num_patches = self.fusion_function.get_output_shape(1, self.time_mixer.num_patch, dim=1)
        self.fusion_mixer = modules.get_block_by_name(**m.multimodal, num_patches=num_patches, dropout=dropout)
        self.classifier_static = torch.nn.Linear(m.static.output_dim, m.classification.num_classes)
        self.classifier_time = torch.nn.Linear(m.time.hidden_dim, m.classification.num_classes)
        self.classifier_fusion = modules.get_classifier_by_name(**m.classification)
        self.fusion_loss_weight = model_cfg.get('fusion_loss_weight', 1.0 / 3)

    def head_weights(self):
        ow = (1 - self.fusion_loss_weight) / 2     # no `* 3` here, unlike AV-MNIST (reference mimic.py:116-121)
        return ow, ow, self.fusion_loss_weight

    def shared_step(self, batch, mode='train', **kwargs):
        static, time, labels = batch
        # the two modality encoders are independent: two streams (models/base.py _two_branches), the short MLP beside the mixer
        time_tokens, static_feat = self._two_branches(lambda: self.time_mixer(time),            # [B, 24, 64]
                                                      lambda: self.static_extractor(static))    # [B, 64]
        fused = self.fusion_mixer(self.fusion_function(static_feat.unsqueeze(1), time_tokens))
        cf = self.classifier_fusion.classifer
        losses, logits, _ = F.heads_loss(
            [static_feat.unsqueeze(1), time_tokens, fused],
            [self.classifier_static.weight, self.classifier_time.weight, cf.weight],
            [self.classifier_static.bias, self.classifier_time.bias, cf.bias],
            labels, self.head_weights(), loss_kind=0)
        return {'preds': torch.softmax(logits[2], dim=1), 'preds_static': torch.softmax(logits[0], dim=1),
                'preds_time': torch.softmax(logits[1], dim=1), 'labels': labels.long(), 'loss': losses[0],
                'loss_fusion': losses[3], 'loss_static': losses[1], 'loss_time': losses[2], 'logits': logits[2],
                'logits_static': logits[0], 'logits_time': logits[1]}
