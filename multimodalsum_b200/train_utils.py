"""Train-step glue of the reference around the fused step.

  get_optimizer / get_scheduler       src/train_utils.py:49-63       (re-exported from .optim, incl. quirk Q1)
  train_step                          src/multimodal_train.py:357-364 forward, zero_grad, backward (+ bucketed all-reduce),
                                      clip_grad_norm_, AdamW.step, scheduler.step
  save_checkpoint                     src/train_utils.py:79-97       pytorch_model.bin (+ training_state.bin), `save_option`
                                      'whole' | 'text' | 'img' | 'table' selects the sub-module whose state_dict is written
The files interchange with the reference: `state_dict` keys / shapes are the reference's, the optimizer state
uses transformers-AdamW names.
"""
import os

import torch

from .optim import FusedAdamW, LinearWarmupSchedule, get_optimizer, get_scheduler  # noqa: F401


def train_step(model, optimizer, scheduler, inputs, reducer=None):
    """One optimizer step exactly in the reference's order.  `optimizer` is a FusedAdamW built with `max_grad_norm`
    (the clip is folded into the update) or any torch optimizer (then `clip_grad_norm_` runs as in the reference)."""
    loss = model(*inputs)[0]
    optimizer.zero_grad()
    loss.backward()                      # a GradAllReducer (dp.py) hooks the engine and overlaps the all-reduce with this call
    if not isinstance(optimizer, FusedAdamW):
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss


def save_checkpoint(model, optimizer, scheduler, epoch, ckpt_dir, save_option="whole"):
    """src/train_utils.py:79-97 (rank 0 only in the reference): writes `pytorch_model.bin` and `training_state.bin`."""
    os.makedirs(ckpt_dir, exist_ok=True)
    save_model = getattr(model, "module", model)
    if save_option == "text":
        save_model = save_model.bart_model
    elif save_option == "img":
        save_model = save_model.img_encoder
    elif save_option == "table":
        save_model = save_model.table_encoder
    sd = {k: v.detach().cpu().clone() for k, v in save_model.state_dict().items()}
    torch.save(sd, os.path.join(ckpt_dir, "pytorch_model.bin"))
    state = {"epoch": epoch, "optimizer": optimizer.state_dict() if optimizer is not None else None,
             "scheduler": scheduler.state_dict() if scheduler is not None else None}
    torch.save(state, os.path.join(ckpt_dir, "training_state.bin"))
    return ckpt_dir
