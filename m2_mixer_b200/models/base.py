"""Lightning-free stand-in for the reference's AbstractTrainTestModule (modules/train_test_module.py:14-175): the
same hooks a Trainer calls (training_step / validation_step / test_step / configure_optimizers) over ``shared_step``,
without wandb / torchmetrics (neither is on the hot path; SURVEY 2 marks them BOUNDARY)."""
from __future__ import annotations

from typing import Any, Dict

import torch
from torch import nn

from ..config import Cfg, wrap
from ..optim import FusedAdam


class TrainTestModule(nn.Module):
    def __init__(self, optimizer_cfg=None, **kwargs):
        super().__init__()
        self.optimizer_cfg = wrap(dict(optimizer_cfg or {}))
        self.scheduler_patience = self.optimizer_cfg.pop("scheduler_patience", 5)
        self.current_epoch = 0
        self.checkpoint_path = None

    # --- the Lightning-facing surface -------------------------------------------------------------------------
    def shared_step(self, batch, **kwargs) -> Dict[str, Any]:  # pragma: no cover
        raise NotImplementedError

    def training_step(self, batch, batch_idx=0):
        return self.shared_step(batch, mode="train")["loss"]

    @torch.no_grad()
    def validation_step(self, batch, batch_idx=0):
        return self.shared_step(batch, mode="val")

    @torch.no_grad()
    def test_step(self, batch, batch_idx=0):
        return self.shared_step(batch, mode="test")

    def configure_optimizers(self):
        """Adam over the trainable parameters + ReduceLROnPlateau (reference models/avmnist.py:413-422)."""
        opt = FusedAdam([p for p in self.parameters() if p.requires_grad], **dict(self.optimizer_cfg))
        sched = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, patience=self.scheduler_patience)
        return {"optimizer": opt, "lr_scheduler": sched, "monitor": "val_loss"}

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict: bool = True, **kwargs):
        """Reference checkpoints carry only ``state_dict`` (no hparams): ctor kwargs come from the YAML."""
        model = cls(**kwargs)
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        model.load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt, strict=strict)
        model.checkpoint_path = checkpoint_path
        return model

    def set_precision(self, p: str) -> "TrainTestModule":
        for m in self.modules():
            if hasattr(m, "precision") and m is not self:
                try:
                    m.precision = p
                except AttributeError:
                    pass
        return self
