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
