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


GEN_CFG = dict(encoder_layers=2, decoder_layers=2, ffn_dim=256, vocab_size=512, max_position_embeddings=256, dropout=0.0,
               dataset="yelp")
GEN_CASES = {
    # name: (cfg kwargs, state-dict kwargs, batch kwargs, generate kwargs) — BASELINE config 5 semantics at toy size
    "gen_small_yelp_s128": (GEN_CFG, dict(seed=21, gates_open=True), dict(B=3, seed=31, n_reviews=3, max_imgs=2, seq_len=128, len_range=(68, 118)),
                            dict(num_beams=4, max_length=24, length_penalty=1.0, no_repeat_ngram_size=3, early_stopping=True)),
    # tie-free variant: a wide random final_logits_bias (seed 77, std 2) separates the candidates by far more than bf16 noise,
    # so the CUDA path must reproduce the ids exactly
    "gen_small_yelp_biased": (GEN_CFG, dict(seed=21, gates_open=True, logits_bias_std=2.0), dict(B=3, seed=33, n_reviews=3, max_imgs=2, seq_len=150, len_range=(90, 140)),
                              dict(num_beams=4, max_length=24, length_penalty=1.0, no_repeat_ngram_size=3, early_stopping=True)),
    "gen_small_yelp_s150": (GEN_CFG, dict(seed=21, gates_open=True), dict(B=3, seed=32, n_reviews=3, max_imgs=2, seq_len=150, len_range=(90, 140)),
                            dict(num_beams=4, max_length=24, length_penalty=1.0, no_repeat_ngram_size=3, early_stopping=True)),
}


def make_generation_goldens():
    for name, (ck, sk, bk, gk) in GEN_CASES.items():
        t0 = time.time()
        cfg = ModelConfig(**ck)
        sd = make_state_dict(cfg, **sk)
        bk2 = dict(bk)
        batch = make_batch(cfg, bk2.pop("B"), **bk2)
        out = RH.reference_generate(cfg, sd, batch, **gk)
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            case=json.dumps(dict(name=name, cfg=ck, sd=sk, batch=bk, gen=gk, torch=torch.__version__, ref_dtype="float32")),
                            tokens=out.numpy())
        print("%s: tokens %s, %.1fs" % (name, tuple(out.shape), time.time() - t0), flush=True)


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
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "gen"):
        make_generation_goldens()
    if which != "gen":
        main(which)
