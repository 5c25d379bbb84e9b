"""CPU: the optimizer / schedule host logic against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py opt: src/train_utils.py:49-63 get_optimizer / get_scheduler, the vendored transformers-3.0.2
AdamW src/transformer/optimization.py:168-267, torch clip_grad_norm_).  The torch restatement below is the checker the
GPU test (tests/test_optim_gpu.py) does NOT need — it compares the fused kernel with the golden directly — but it shows on
CPU that the golden is reachable from the documented formula, including quirk Q1 (the no-decay group is empty)."""
import math
import types

import torch

from golden_util import check_params_against_optimizer_golden, load_optimizer_golden
from multimodalsum_b200.optim import LinearWarmupSchedule, get_scheduler


def test_linear_warmup_schedule_matches_reference_table():
    gold = load_optimizer_golden()
    c = gold["case"]
    opt = types.SimpleNamespace(param_groups=[{"lr": c["lr"]}, {"lr": c["lr"]}])
    sched = get_scheduler(types.SimpleNamespace(num_epochs=c["num_epochs"], warmup_ratio=c["warmup_ratio"]), c["t_epoch"], opt)
    assert isinstance(sched, LinearWarmupSchedule)
    got = [opt.param_groups[0]["lr"]]
    for _ in range(len(gold["sched_lrs"]) - 1):
        sched.step()
        got.append(opt.param_groups[1]["lr"])
    assert got == gold["sched_lrs"], (got, gold["sched_lrs"])
    # state round trip
    s2 = LinearWarmupSchedule(types.SimpleNamespace(param_groups=[{"lr": c["lr"]}]), 2, 4)
    s2.load_state_dict(dict(sched.state_dict(), last_epoch=1))
    assert s2.get_last_lr() == [gold["sched_lrs"][1]]


def test_documented_adamw_formula_reproduces_reference_golden():
    gold = load_optimizer_golden()
    c = gold["case"]
    params = {n: gold["sd"][n].clone() for n in gold["names"]}
    m = {n: torch.zeros_like(p) for n, p in params.items()}
    v = {n: torch.zeros_like(p) for n, p in params.items()}
    gnorm = math.sqrt(sum(float((g.double() ** 2).sum()) for g in gold["grads"].values()))
    assert abs(gnorm - gold["gnorms"][0]) <= 5e-5 * gnorm      # torch's clip_grad_norm_ accumulates the norm in fp32
    clip = min(1.0, c["max_grad_norm"] / (gnorm + 1e-6))
    b1, b2, eps = 0.9, 0.999, 1e-6
    for t, lr in enumerate(gold["lrs"], start=1):
        for n, p in params.items():
            if any(nd in n for nd in c["no_decay"]):
                continue                                   # quirk Q1: never updated by the reference
            g = gold["grads"][n] * clip
            m[n].mul_(b1).add_(g, alpha=1 - b1)
            v[n].mul_(b2).addcmul_(g, g, value=1 - b2)
            step_size = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
            p.addcdiv_(m[n], v[n].sqrt().add_(eps), value=-step_size)
            p.add_(p, alpha=-lr * 0.01)
    assert not check_params_against_optimizer_golden(gold, params, rtol=1e-4)
