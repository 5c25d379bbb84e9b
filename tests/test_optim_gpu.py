"""Fused clip + AdamW vs a torch restatement of clip_grad_norm_ + transformers-3.0.2 AdamW
(src/transformer/optimization.py:208-267) on the real parameter arena of a small model."""
import math

import pytest
import torch

from golden_util import load_golden

pytestmark = pytest.mark.gpu


def _ref_adamw(p, g, m, v, step, lr, wd, b1=0.9, b2=0.999, eps=1e-6):
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    denom = v.sqrt().add_(eps)
    step_size = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if wd > 0:
        p.add_(p, alpha=-lr * wd)


def test_fused_adamw_matches_reference_semantics_incl_quirk_q1():
    from test_step_gpu import _run_cuda_step
    from multimodalsum_b200.optim import get_optimizer
    gold = load_golden("small_yelp")
    _, grads, model = _run_cuda_step(gold)
    eng = model.engine
    no_decay = ["bias", "LayerNorm.weight"]
    # a GENERATOR, as in src/multimodal_train.py:462 -> the no-decay group is empty (quirk Q1)
    opt = get_optimizer(eng, 1e-3, no_decay, model.named_parameters(), None, max_grad_norm=1.0)
    assert len(opt.param_groups[1]["params"]) == 0
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    gnorm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    clip = min(1.0, 1.0 / (gnorm + 1e-6))
    state = {n: (torch.zeros_like(p), torch.zeros_like(p)) for n, p in before.items()}
    for step in (1, 2):
        opt.step()
        torch.cuda.synchronize()
        assert abs(opt.grad_norm().item() - gnorm) <= 1e-4 * gnorm
        for n, p in before.items():
            if any(nd in n for nd in no_decay):
                continue                      # never updated by the reference
            _ref_adamw(p, grads[n] * clip, state[n][0], state[n][1], step, 1e-3, 0.01)
    for n, p in model.named_parameters():
        assert torch.allclose(p.detach(), before[n], rtol=2e-5, atol=2e-7), n
    # biases are untouched (Q1) and the bf16 compute copy follows the masters
    b = "bart_model.model.encoder.layers.0.fc1.bias"
    assert torch.equal(dict(model.named_parameters())[b].detach(), gold["sd"][b].cuda())
    w = "bart_model.model.encoder.layers.0.fc1.weight"
    assert torch.equal(eng.w16(w), eng.w32(w).to(torch.bfloat16))
    # the next forward uses the updated weights without a separate cast pass
    loss2 = model(*_inputs(gold))[0]
    assert torch.isfinite(loss2) and abs(loss2.item() - gold["loss"]) > 1e-5


def _inputs(gold):
    b = gold["batch"].to("cuda")
    return b.reviews, b.reviews_mask, b.reviews_rating, b.field, b.field_value, b.img, b.img_mask
