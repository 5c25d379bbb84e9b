"""Fused clip + AdamW + linear warm-up on the GPU against the golden produced by the UNMODIFIED reference
(tests/golden/adamw_small.npz: src/train_utils.py:49-63 get_optimizer / get_scheduler incl. quirk Q1, the vendored
transformers-3.0.2 AdamW src/transformer/optimization.py:168-267, torch clip_grad_norm_, driven in the order of
src/multimodal_train.py:359-364) on synthetic gradients that are rebuilt from seeds on both sides."""
import types

import pytest
import torch

from golden_util import check_params_against_optimizer_golden, load_optimizer_golden

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_reference_golden_incl_quirk_q1_and_schedule():
    from multimodalsum_b200.modules import MultimodalSum, YelpTableEncoder
    from multimodalsum_b200.optim import get_optimizer, get_scheduler
    gold = load_optimizer_golden()
    c = gold["case"]
    model = MultimodalSum(TableEncoder=YelpTableEncoder, config=gold["cfg"])
    model.load_state_dict(gold["sd"], strict=False)
    model = model.cuda()
    eng = model._ensure_engine(torch.device("cuda"))
    # a GENERATOR, as in src/multimodal_train.py:462 -> the no-decay group is empty (quirk Q1)
    opt = get_optimizer(eng, c["lr"], c["no_decay"], model.named_parameters(), None, max_grad_norm=c["max_grad_norm"])
    assert len(opt.param_groups[1]["params"]) == 0
    sched = get_scheduler(types.SimpleNamespace(num_epochs=c["num_epochs"], warmup_ratio=c["warmup_ratio"]), c["t_epoch"], opt)
    for step in range(c["steps"]):
        for n in eng.names:
            eng.g32(n).copy_(gold["grads"][n])
        assert abs(opt.param_groups[0]["lr"] - gold["lrs"][step]) <= 1e-12
        opt.step()
        sched.step()
        torch.cuda.synchronize()
        assert abs(opt.grad_norm().item() - gold["gnorms"][step]) <= 5e-5 * gold["gnorms"][step]
    params = dict(model.named_parameters())
    bad = check_params_against_optimizer_golden(gold, params, rtol=1e-4)
    assert not bad, bad[:5]
    # biases are untouched (Q1) and the bf16 compute copy follows the masters without a separate cast pass
    b = "bart_model.model.encoder.layers.0.fc1.bias"
    assert torch.equal(params[b].detach().cpu(), gold["sd"][b])
    w = "bart_model.model.encoder.layers.0.fc1.weight"
    assert torch.equal(eng.w16(w), eng.w32(w).to(torch.bfloat16))
    assert eng._w16_fresh
    # optimizer state round trip
    sd = opt.state_dict()
    opt2 = get_optimizer(eng, c["lr"], c["no_decay"], model.named_parameters(), None, max_grad_norm=c["max_grad_norm"])
    opt2.load_state_dict(sd)
    assert opt2.step_count == c["steps"] and torch.equal(opt2.m, opt.m) and torch.equal(opt2.v, opt.v)
