"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (/root/reference) on the
synthetic weights/inputs of multimodalsum_b200.synth.  Run in the build container only:

    python tests/golden/make_golden.py [small|full|all]

Each .npz holds: the case description (json), the reference's fp32 loss, the L2 norm of every parameter gradient
and the first 64 entries (stride 997) of every gradient.  Weights/inputs are NOT stored: they are rebuilt from the
seeds in the case description (torch CPU generators are deterministic for the pinned torch build).
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from multimodalsum_b200.synth import ModelConfig, make_batch, make_state_dict  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

SMALL = dict(encoder_layers=2, decoder_layers=2, ffn_dim=256, vocab_size=512, max_position_embeddings=128, dropout=0.0)
CASES = {
    # name: (cfg kwargs, state-dict kwargs, batch kwargs)
    "small_yelp": (dict(SMALL, dataset="yelp"), dict(seed=0), dict(B=2, seed=1, n_reviews=4, max_imgs=2)),
    "small_yelp_gates_open": (dict(SMALL, dataset="yelp"), dict(seed=0, gates_open=True), dict(B=3, seed=2, n_reviews=3, max_imgs=3)),
    "small_amazon": (dict(SMALL, dataset="amazon"), dict(seed=3), dict(B=2, seed=4, n_reviews=4)),
    "small_text": (dict(SMALL, dataset="text"), dict(seed=5), dict(B=2, seed=6, n_reviews=3)),
    "full_yelp_b1": (dict(dataset="yelp", dropout=0.0), dict(seed=0), dict(B=1, seed=1, n_reviews=9, n_valid_imgs=4)),
    "full_yelp_b1_gates_open": (dict(dataset="yelp", dropout=0.0), dict(seed=0, gates_open=True), dict(B=1, seed=7, n_reviews=9)),
    "full_text_b1": (dict(dataset="text", dropout=0.0), dict(seed=0), dict(B=1, seed=8, n_reviews=9)),
}


def build_case(name):
    ck, sk, bk = CASES[name]
    cfg = ModelConfig(**ck)
    sd = make_state_dict(cfg, **sk)
    bk = dict(bk)
    batch = make_batch(cfg, bk.pop("B"), **bk)
    return cfg, sd, batch


def sample(g):
    return g.flatten()[::997][:64].double().numpy()


def main(which):
    for name in CASES:
        if which != "all" and not name.startswith(which):
            continue
        t0 = time.time()
        cfg, sd, batch = build_case(name)
        if cfg.dataset == "text":
            loss, grads, _ = RH.reference_text_step(cfg, sd, batch, dtype=torch.float32, label_smoothing=None)
        else:
            loss, grads, _ = RH.reference_step(cfg, sd, batch, dtype=torch.float32, label_smoothing=0.1)
        out = {"case": json.dumps(dict(name=name, cfg=CASES[name][0], sd=CASES[name][1], batch=CASES[name][2],
                                       torch=torch.__version__, ref_dtype="float32")),
               "loss": np.float64(loss.item())}
        names = sorted(grads)
        out["names"] = np.array(names)
        out["norms"] = np.array([grads[n].double().norm().item() for n in names])
        out["samples"] = np.stack([np.pad(sample(grads[n]), (0, 64 - len(sample(grads[n])))) for n in names])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("%s: loss %.9f, %d grads, %.1fs" % (name, loss.item(), len(names), time.time() - t0), flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "all")
