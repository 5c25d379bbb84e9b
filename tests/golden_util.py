"""Helpers shared by the parity tests: rebuild a golden case from its description and compare gradients."""
import json
import os

import numpy as np
import torch

from multimodalsum_b200.synth import ModelConfig, make_batch, make_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    case = json.loads(str(z["case"]))
    cfg = ModelConfig(**case["cfg"])
    sd = make_state_dict(cfg, **case["sd"])
    bk = dict(case["batch"])
    batch = make_batch(cfg, bk.pop("B"), **bk)
    names = [str(n) for n in z["names"]]
    return dict(cfg=cfg, sd=sd, batch=batch, loss=float(z["loss"]), names=names,
                norms=dict(zip(names, z["norms"].tolist())), samples=dict(zip(names, z["samples"])),
                label_smoothing=None if cfg.dataset == "text" else 0.1)


def load_optimizer_golden(name="adamw_small"):
    """Parameters after `steps` rounds of clip_grad_norm_ + AdamW + linear warm-up as the UNMODIFIED reference runs them
    (tests/golden/make_golden.py opt)."""
    from multimodalsum_b200.synth import make_grads
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    case = json.loads(str(z["case"]))
    cfg = ModelConfig(**case["cfg"])
    names = [str(n) for n in z["names"]]
    return dict(case=case, cfg=cfg, sd=make_state_dict(cfg, **case["sd"]), grads=make_grads(cfg, **case["grads"]), names=names,
                lrs=z["lrs"].tolist(), gnorms=z["gnorms"].tolist(), norms=dict(zip(names, z["norms"].tolist())),
                delta_norms=dict(zip(names, z["delta_norms"].tolist())), samples=dict(zip(names, z["samples"])),
                sched_lrs=z["sched_lrs"].tolist())


def check_params_against_optimizer_golden(gold, params, rtol=2e-5):
    """params: {name: tensor} after the golden's steps.  The UPDATE (p - p0) must match to rtol of its own norm; sampled
    entries of p must match to rtol of the update's RMS."""
    bad = []
    for n in gold["names"]:
        p = params[n].detach().double().cpu()
        p0 = gold["sd"][n].double()
        dn = (p - p0).norm().item()
        ref_dn = gold["delta_norms"][n]
        smp = sample(p)
        rms = max(ref_dn, 1e-12) / max(p.numel(), 1) ** 0.5
        e_smp = float(np.abs(smp - gold["samples"][n]).max())
        if abs(dn - ref_dn) > rtol * max(ref_dn, 1e-12) + 1e-12 or e_smp > 50 * rtol * rms + 1e-9:
            bad.append((n, dn, ref_dn, e_smp, rms))
    return bad


def sample(g):
    s = g.detach().flatten()[::997][:64].double().cpu().numpy()
    return np.pad(s, (0, 64 - len(s)))


def weight_sibling(name):
    return name[:-len(".bias")] + ".weight" if name.endswith(".bias") else name


def compare_grads(gold, grads, rel_tol, report=None):
    """Per-tensor check: |‖g‖-‖g_ref‖| and the sampled entries within rel_tol of the tensor's own norm.
    k_proj.bias gradients are identically zero in exact arithmetic (softmax shift invariance); they are checked
    against the norm of the sibling weight gradient instead."""
    bad = []
    for n in gold["names"]:
        assert n in grads, "missing gradient for %s" % n
        g = grads[n]
        ref_norm = gold["norms"][n]
        scale = ref_norm
        if n.endswith("k_proj.bias"):
            scale = max(gold["norms"][weight_sibling(n)], ref_norm)
        norm = g.double().norm().item()
        e_norm = abs(norm - ref_norm) / max(scale, 1e-30)
        smp = sample(g)
        # sampled entries: compare relative to the RMS entry size implied by the norm
        rms = scale / max(g.numel(), 1) ** 0.5
        e_smp = float(np.abs(smp - gold["samples"][n]).max()) / max(rms, 1e-30)
        if report is not None:
            report.append((n, e_norm, e_smp))
        if e_norm > rel_tol or e_smp > 40 * rel_tol:
            bad.append((n, e_norm, e_smp, ref_norm))
    return bad
