"""MMIMDBMixerMultiLoss (reference models/mmimdb.py:21-147): image + text encoders, 23-label multilabel,
BCEWithLogitsLoss(pos_weight) x 3 summed.  The text encoder is whatever ``modalities.text.block_type`` names
(MLPMixer in the shipped cfg, PNLPMixer in the synthetic C4 config of SURVEY 8d)."""
from __future__ import annotations

import torch

from .. import functional as F
from .. import modules
from ..config import wrap
from .base import TrainTestModule


class MMIMDBMixerMultiLoss(TrainTestModule):
    def __init__(self, model_cfg, optimizer_cfg=None, **kwargs):
        super().__init__(optimizer_cfg, **kwargs)
        model_cfg = wrap(model_cfg)
        self.modalities_freezed = False
        m = model_cfg.modalities
        dropout = model_cfg.get('dropout', 0.0)
        self.image_mixer = modules.get_block_by_name(**m.image, dropout=dropout)
        self.text_mixer = modules.get_block_by_name(**m.text, dropout=dropout)
        self.fusion_function = modules.get_fusion_by_name(**m.multimodal)
        num_patches = self.fusion_function.get_output_shape(self.image_mixer.num_patch, self.text_mixer.num_patch, dim=1)
        self.fusion_mixer = modules.get_block_by_name(**m.multimodal, num_patches=num_patches, dropout=dropout)
        self.classifier_image = torch.nn.Linear(m.image.hidden_dim, m.classification.num_classes)
        self.classifier_text = torch.nn.Linear(m.text.hidden_dim, m.classification.num_classes)
        self.classifier_fusion = modules.get_classifier_by_name(**m.classification)
        pw = model_cfg.get('pos_weight', None)
        self.register_buffer('pos_weight', None if pw is None else torch.tensor(list(pw), dtype=torch.float32),
                             persistent=False)

    def shared_step(self, batch, **kwargs):
        image, text, labels = batch['image'], batch['text'], batch['label']
        toks, slices, _ = self._encode_and_fuse(self.image_mixer, image, self.text_mixer, text, self.fusion_function,
                                                self.fusion_mixer)
        hw = (0.0, 0.0, 1.0) if (self.modalities_freezed and kwargs.get('mode') == 'train') else (1.0, 1.0, 1.0)
        cf = self.classifier_fusion.classifer
        losses, logits, preds = F.heads_loss(
            toks,
            [self.classifier_image.weight, self.classifier_text.weight, cf.weight],
            [self.classifier_image.bias, self.classifier_text.bias, cf.bias],
            labels, hw, loss_kind=1, pos_weight=self.pos_weight, slices=slices)
        return {'preds': preds[2], 'preds_image': preds[0], 'preds_text': preds[1], 'labels': labels, 'loss': losses[0],
                'loss_image': losses[1], 'loss_text': losses[2], 'loss_fusion': losses[3], 'image_logits': logits[0],
                'text_logits': logits[1], 'logits': logits[2]}
