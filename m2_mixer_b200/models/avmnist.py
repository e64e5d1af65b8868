"""AVMnistMixerMultiLoss: same class name (= cfg ``model.type``), ctor ``(model_cfg, optimizer_cfg)``, sub-module
names (= state-dict prefixes) and ``shared_step`` return dict as the reference (models/avmnist.py:166-312)."""
from __future__ import annotations

import numpy as np
import torch

from .. import functional as F
from .. import modules
from ..config import wrap
from .base import TrainTestModule


class AVMnistMixerMultiLoss(TrainTestModule):
    def __init__(self, model_cfg, optimizer_cfg=None, **kwargs):
        super().__init__(optimizer_cfg, **kwargs)
        model_cfg = wrap(model_cfg)
        self.modalities_freezed = False
        self.mute = model_cfg.get('mute', None)
        self.freeze_modalities_on_epoch = model_cfg.get('freeze_modalities_on_epoch', None)
        self.random_modality_muting_on_freeze = model_cfg.get('random_modality_muting_on_freeze', False)
        self.muting_probs = model_cfg.get('muting_probs', None)
        m = model_cfg.modalities
        dropout = model_cfg.get('dropout', 0.0)
        self.image_mixer = modules.get_block_by_name(**m.image, dropout=dropout)
        self.audio_mixer = modules.get_block_by_name(**m.audio, dropout=dropout)
        self.fusion_function = modules.get_fusion_by_name(**m.multimodal)
        num_patches = self.fusion_function.get_output_shape(self.image_mixer.num_patch, self.audio_mixer.num_patch, dim=1)
        self.fusion_mixer = modules.get_block_by_name(**m.multimodal, num_patches=num_patches, dropout=dropout)
        self.classifier_image = torch.nn.Linear(m.image.hidden_dim, m.classification.num_classes)
        self.classifier_audio = torch.nn.Linear(m.audio.hidden_dim, m.classification.num_classes)
        self.classifier_fusion = modules.get_classifier_by_name(**m.classification)
        self.fusion_loss_weight = model_cfg.get('fusion_loss_weight', 1.0 / 3)
        self.fusion_loss_change = model_cfg.get('fusion_loss_change', 0)
        self.loss_change_epoch = model_cfg.get('loss_change_epoch', 0)

    def head_weights(self, mode=None):
        """(image, audio, fusion) multipliers of the summed loss - runtime scalars (reference :290-293, :338-339)."""
        if self.modalities_freezed and mode == 'train':
            return 0.0, 0.0, 1.0
        ow = (1 - self.fusion_loss_weight) / 2
        return ow * 3, ow * 3, self.fusion_loss_weight * 3

    def shared_step(self, batch, **kwargs):
        image, audio, labels = batch['image'], batch['audio'], batch['label']
        mode = kwargs.get('mode', None)
        if mode == 'train':
            if self.freeze_modalities_on_epoch is not None and self.current_epoch == self.freeze_modalities_on_epoch \
                    and not self.modalities_freezed:
                self._freeze_modalities()
            if self.random_modality_muting_on_freeze and self.current_epoch >= self.freeze_modalities_on_epoch:
                self.mute = np.random.choice(['image', 'audio', 'multimodal'],
                                             p=[self.muting_probs['image'], self.muting_probs['audio'],
                                                self.muting_probs['multimodal']])
            if self.mute == 'image':
                image = torch.zeros_like(image)
            elif self.mute == 'audio':
                audio = torch.zeros_like(audio)

        toks, slices, _ = self._encode_and_fuse(self.image_mixer, image, self.audio_mixer, audio, self.fusion_function,
                                                self.fusion_mixer)

        # three mean-pool + Linear heads, three cross entropies and their weighted sum: one kernel
        cf = self.classifier_fusion.classifer
        losses, logits, preds = F.heads_loss(
            toks,
            [self.classifier_image.weight, self.classifier_audio.weight, cf.weight],
            [self.classifier_image.bias, self.classifier_audio.bias, cf.bias],
            labels, self.head_weights(mode), loss_kind=0, slices=slices)
        return {'preds': preds[2], 'preds_image': preds[0], 'preds_audio': preds[1], 'labels': labels,
                'loss': losses[0], 'loss_image': losses[1], 'loss_audio': losses[2], 'loss_fusion': losses[3],
                'image_logits': logits[0], 'audio_logits': logits[1], 'logits': logits[2]}

    def _freeze_modalities(self):
        for mod in (self.image_mixer, self.audio_mixer, self.classifier_image, self.classifier_audio):
            for p in mod.parameters():
                p.requires_grad = False
        self.modalities_freezed = True

    def on_train_epoch_end(self):
        if self.current_epoch >= self.loss_change_epoch:
            self.fusion_loss_weight = min(self.fusion_loss_weight + self.fusion_loss_change, 1.0)
        self.current_epoch += 1

    def test_epoch_end(self, outputs, save_dir=None):
        """Concatenate the per-batch ``test_step`` dicts and write ``test_preds.pt`` next to the checkpoint, as the reference
        does (models/avmnist.py:382-398: same keys, same file name).  ``save_dir`` replaces the Lightning logger path the
        reference falls back to when no checkpoint path is known."""
        import os
        keys = ("preds", "preds_image", "preds_audio", "labels", "image_logits", "audio_logits", "logits")
        cat = {k: torch.cat([o[k].detach().cpu() for o in outputs]) for k in keys}
        if save_dir is None:
            if self.checkpoint_path is None:
                raise ValueError("test_epoch_end needs checkpoint_path (load_from_checkpoint sets it) or save_dir")
            save_dir = os.path.dirname(self.checkpoint_path)
        os.makedirs(save_dir, exist_ok=True)
        out = os.path.join(save_dir, "test_preds.pt")
        torch.save(cat, out)
        return out
