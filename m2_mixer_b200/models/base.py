"""Lightning-free stand-in for the reference's AbstractTrainTestModule (modules/train_test_module.py:14-175): the
same hooks a Trainer calls (training_step / validation_step / test_step / configure_optimizers) over ``shared_step``,
without wandb / torchmetrics (neither is on the hot path; SURVEY 2 marks them BOUNDARY)."""
from __future__ import annotations

from typing import Any, Dict

import torch
from torch import nn

from ..config import Cfg, wrap
from ..optim import FusedAdam


_SIDE_STREAMS: Dict[Any, "torch.cuda.Stream"] = {}


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

    # --- two encoders -> ConcatFusion -> fusion mixer -> three heads ----------------------------------------------
    def _encode_and_fuse(self, mixer_a, xa, mixer_b, xb, fusion_function, fusion_mixer):
        """(tokens per head, token slices, fused tokens) of the reference's
        ``fusion_mixer(fusion_function(mixer_a(xa), mixer_b(xb)))`` (models/avmnist.py:262-266, models/mmimdb.py:100-104).
        When the fusion is ConcatFusion(dim=1) over two Mixer stacks of one width, the closing LayerNorms write straight
        into the fused-token buffer (no concat copy, no split copies in the backward) and the per-modality heads pool their
        slices of that buffer in place."""
        from .. import functional as F
        from ..modules.fusion import ConcatFusion
        from ..modules.mixer import _Stack
        if (isinstance(fusion_function, ConcatFusion) and fusion_function.dim == 1 and isinstance(mixer_a, _Stack)
                and isinstance(mixer_b, _Stack)
                and mixer_a.layer_norm.weight.shape == mixer_b.layer_norm.weight.shape):
            fa, fb = self._two_branches(lambda: mixer_a.forward_features(xa), lambda: mixer_b.forward_features(xb))
            la, lb = mixer_a.layer_norm, mixer_b.layer_norm
            cat = F.layer_norm_concat([fa, fb], [la.weight, lb.weight], [la.bias, lb.bias])
            fused = fusion_mixer(cat)
            na, nb = fa.shape[1], fb.shape[1]
            return [cat, cat, fused], [(0, na), (na, nb), (0, fused.shape[1])], fused
        ta, tb = self._two_branches(lambda: mixer_a(xa), lambda: mixer_b(xb))
        fused = fusion_mixer(fusion_function(ta, tb))
        return [ta, tb, fused], None, fused

    def _two_branches(self, fn_a, fn_b):
        """Run the two encoders on two streams (autograd then runs their backwards on the same two streams).  Every kernel of
        an M2-Mixer-B encoder block is ONE persistent CTA per SM on 128 (channel mixing) or 144 (weight gradients) of the 148
        SMs: alone, each leaves 3-14 % of the GPU idle for its whole duration, and the next kernel of the same branch cannot
        start before its last CTA retires.  The two branches are independent until the fusion, so the block scheduler fills
        one's idle SMs and tails with the other's CTAs - also inside a captured graph (fork / join become graph edges).
        M2B200_BRANCH_STREAMS=0 keeps one stream (A/B measurements)."""
        import os
        from .. import functional as F
        if not torch.cuda.is_available() or os.environ.get("M2B200_BRANCH_STREAMS", "1") == "0":
            return fn_a(), fn_b()
        cur = torch.cuda.current_stream()
        side = _SIDE_STREAMS.get(cur.device)      # one per device, shared by every model (not module state: deepcopy / pickle)
        if side is None:
            side = _SIDE_STREAMS[cur.device] = torch.cuda.Stream(device=cur.device)
        F.COMPUTE_STREAMS[:] = [cur, side]
        F.ensure_bf16_weights_fresh()              # on the caller's stream, before the fork (see functional.py)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            b = fn_b()
        a = fn_a()
        cur.wait_stream(side)
        if not torch.cuda.is_current_stream_capturing():
            for t in (b if isinstance(b, (tuple, list)) else (b,)):
                if torch.is_tensor(t):
                    t.record_stream(cur)           # allocated on the side stream, consumed on the caller's
        return a, b

    def set_precision(self, p: str) -> "TrainTestModule":
        for m in self.modules():
            if hasattr(m, "precision") and m is not self:
                try:
                    m.precision = p
                except AttributeError:
                    pass
        return self
