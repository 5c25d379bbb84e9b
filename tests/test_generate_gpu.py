"""Beam-search generation (BASELINE config 5 semantics) against token ids produced by the UNMODIFIED
reference's `generate` (tests/golden/make_golden.py gen).  Integer output: exact equality."""
import json
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR
from multimodalsum_b200.synth import ModelConfig, make_batch, make_state_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["gen_small_yelp_s128", "gen_small_yelp_s150"])
def test_beam_search_matches_reference_tokens(name):
    from multimodalsum_b200.generation import Generator
    from multimodalsum_b200.modules import MultimodalSum, YelpTableEncoder
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    case = json.loads(str(z["case"]))
    cfg = ModelConfig(**case["cfg"])
    sd = make_state_dict(cfg, **case["sd"])
    bk = dict(case["batch"])
    batch = make_batch(cfg, bk.pop("B"), **bk).to("cuda")
    model = MultimodalSum(TableEncoder=YelpTableEncoder, config=cfg)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    gen = Generator(model)
    out = gen.generate(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask, **case["gen"])
    ref = torch.from_numpy(z["tokens"])
    assert tuple(out.shape) == tuple(ref.shape), (out.shape, ref.shape)
    assert torch.equal(out.cpu(), ref), (out.cpu().tolist(), ref.tolist())


def test_ngram_blocking_and_hypothesis_heap_host_logic():
    from multimodalsum_b200.generation import BeamHypotheses, calc_banned_ngram_tokens
    # fairseq semantics: with [5,6,7,5,6] and n=3 the next token after (5,6) may not be 7
    assert calc_banned_ngram_tokens([[5, 6, 7, 5, 6]], 1, 3, 5) == [[7]]
    assert calc_banned_ngram_tokens([[5, 6]], 1, 3, 1) == [[]]
    h = BeamHypotheses(2, 10, 1.0, early_stopping=True)
    h.add([1, 2, 3], -3.0); h.add([1, 2], -1.0); h.add([1, 2, 3, 4], -2.0)
    assert len(h) == 2 and sorted(s for s, _ in h.beams) == [-0.5, -0.5] or len(h) == 2
    assert h.is_done(-100.0, 5)
