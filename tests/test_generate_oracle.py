"""CPU: the beam-search host logic of the product (multimodalsum_b200.generation.beam_search — a restatement of
_generate_beam_search, modeling_multimodalsum.py:2803-3067) fed with the oracle's fp32 logits must reproduce, token for
token, what the UNMODIFIED reference's `generate` produced (tests/golden/gen_*.npz)."""
import json
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR
from multimodalsum_b200.generation import beam_search
from multimodalsum_b200.synth import ModelConfig, make_batch, make_state_dict
from oracle import mmsum_oracle as OR


def load_gen_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    case = json.loads(str(z["case"]))
    cfg = ModelConfig(**case["cfg"])
    sd = make_state_dict(cfg, **case["sd"])
    bk = dict(case["batch"])
    batch = make_batch(cfg, bk.pop("B"), **bk)
    return cfg, sd, batch, case["gen"], torch.from_numpy(z["tokens"])


def oracle_params(sd, device="cpu"):
    p = {k: v.to(device) for k, v in sd.items()}
    return p


@pytest.mark.parametrize("name", ["gen_small_yelp_s128", "gen_small_yelp_s150", "gen_small_yelp_biased"])
def test_beam_search_logic_reproduces_reference_generate(name):
    cfg, sd, batch, gk, ref = load_gen_case(name)
    fn = OR.generation_logits_fn(oracle_params(sd), cfg, batch, gk["num_beams"])
    out = beam_search(fn, batch.reviews.shape[0], cfg.vocab_size, torch.device("cpu"), pad=cfg.pad_token_id, bos=cfg.bos_token_id,
                      eos=cfg.eos_token_id, **gk)
    assert torch.equal(out, ref), (out.tolist(), ref.tolist())
