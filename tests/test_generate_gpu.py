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
    """`step_logits` (one token per hypothesis on the decode-shaped kernels, slot table permuted by `reorder_cache` after
    every step) against `last_logits` (whole prefix recomputed in 128-row frames on the training kernels) on the same token
    histories, 4 beams per business, including beam permutations inside a business: two independent bf16 kernel families,
    each within 0.05 nats of the fp32 oracle -> log-probabilities agree to 4e-2 nats and the arg-max is identical wherever
    the top-2 margin exceeds 0.1 nats."""
    gen, cfg, sd, batch, gk, ref = _setup("gen_small_yelp_s128")
    beams = 4
    B = batch.reviews.shape[0]
    N = B * beams
    st_c = gen.encode(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask, beams)
    st_r = gen.encode(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask, beams)
    _cached_vs_recompute(gen, cfg, st_c, st_r, B, beams, steps=10, tol=4e-2)


def _cached_vs_recompute(gen, cfg, st_c, st_r, B, beams, steps, tol):
    N = B * beams
    rd = torch.zeros(N, device="cuda")
    g = torch.Generator().manual_seed(5)
    ids = torch.full((N, 1), cfg.eos_token_id, dtype=torch.long, device="cuda")
    worst = 0.0
    for step in range(steps):
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
    assert worst <= tol, worst


def test_generation_at_config5_shape():
    """BASELINE configs[4] shape: 64 businesses x 4 beams, 8 reviews in 158-token frames (src/test.py:57), 47 table fields,
    10 x 196 image keys, vocabulary 50265, BART-large widths (2 + 2 layers so the fp32 oracle fits the test budget).
    (a) teacher-forced next-token log-probabilities of all 256 hypotheses vs OR.generation_logits_fn (<= 0.05 nats, arg-max
    equal where the oracle's top-2 margin exceeds 0.1 nats); (b) the cached decoder vs prefix recompute under beam
    permutations; (c) generate() runs end to end and returns well-formed ids."""
    from multimodalsum_b200.generation import Generator
    from multimodalsum_b200.modules import MultimodalSum, YelpTableEncoder
    from oracle import mmsum_oracle as OR
    B, beams = 64, 4
    cfg = ModelConfig(dataset="yelp", encoder_layers=2, decoder_layers=2, dropout=0.0)
    sd = make_state_dict(cfg, seed=41, gates_open=True, logits_bias_std=1.0)
    batch = make_batch(cfg, B, seed=42, n_reviews=8, seq_len=158, len_range=(100, 150)).to("cuda")
    model = MultimodalSum(TableEncoder=YelpTableEncoder, config=cfg)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    gen = Generator(model)
    args = (batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask)
    st = gen.encode(*args, beams)
    N = B * beams
    rd = torch.zeros(N, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    p = {k: v.cuda() for k, v in sd.items()}
    ofn = OR.generation_logits_fn(p, cfg, batch, beams)
    g = torch.Generator().manual_seed(9)
    ids = torch.cat([torch.full((N, 1), cfg.eos_token_id), torch.randint(3, cfg.vocab_size, (N, 5), generator=g)], dim=1).cuda()
    worst, checked = 0.0, 0
    for cur in (1, 3, 6):
        lc = torch.log_softmax(gen.last_logits(st, ids[:, :cur].contiguous(), rd).float(), -1)
        lo = torch.log_softmax(ofn(ids[:, :cur].contiguous()), -1)
        worst = max(worst, (lc - lo).abs().max().item())
        top2 = lo.topk(2, dim=-1).values
        tie_free = (top2[:, 0] - top2[:, 1]) > 0.1
        assert torch.equal(lc.argmax(-1)[tie_free], lo.argmax(-1)[tie_free])
        checked += int(tie_free.sum())
    assert worst <= 0.05, worst
    assert checked > 0
    # the incremental decoder (decode-shaped kernels, CUDA-graph replay from the second token) against the oracle as well
    st_i = gen.encode(*args, beams)
    for cur in range(1, 7):
        li = gen.step_logits(st_i, ids[:, :cur].contiguous(), rd)
        if cur in (1, 3, 6):
            lo = torch.log_softmax(ofn(ids[:, :cur].contiguous()), -1)
            assert (torch.log_softmax(li.float(), -1) - lo).abs().max().item() <= 0.05
    del ofn, p, st_i
    torch.cuda.empty_cache()
    st_c, st_r = gen.encode(*args, beams), st
    _cached_vs_recompute(gen, cfg, st_c, st_r, B, beams, steps=6, tol=4e-2)
    out = gen.generate(*args, num_beams=beams, max_length=12, no_repeat_ngram_size=3, early_stopping=True)
    assert gen.last_used_graph                      # decoder step + beam update of a token replayed from one CUDA graph
    out2 = gen.generate(*args, num_beams=beams, max_length=12, no_repeat_ngram_size=3, early_stopping=True)
    ref_out = gen.generate(*args, num_beams=beams, max_length=12, no_repeat_ngram_size=3, early_stopping=True, use_cache=False)

    def agreement(a, b):     # on the common width, padded on the right (a swapped near-tie can move the longest summary by a token)
        W = max(a.shape[1], b.shape[1])
        padw = lambda t: torch.nn.functional.pad(t, (0, W - t.shape[1]), value=cfg.pad_token_id)
        return float((padw(a) == padw(b)).float().mean())

    # second call: same plan, graph replayed from the first token on — bit-identical (the decode cross-attention adds the entity
    # outputs of a business in entity order; when it used shared-memory atomics in completion order, a near-tied candidate
    # swapped in about one run out of ten and this equality failed at that rate)
    assert torch.equal(out, out2)
    # two bf16 kernel families (cached decode vs prefix recompute): near-tied candidates may swap in a few businesses
    same = agreement(out, ref_out)
    assert same > 0.9, (same, tuple(out.shape), tuple(ref_out.shape))
    assert out.shape[0] == B and out.shape[1] <= 12 and (out[:, 0] == cfg.eos_token_id).all() and (out[:, 1] == cfg.bos_token_id).all()


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




@pytest.mark.parametrize("k,V,ngram,min_length", [(4, 50265, 3, 0), (1, 1000, 2, 6), (8, 777, 3, 0), (2, 50265, 0, 0)])
def test_beam_candidate_kernel_matches_tensor_implementation(k, V, ngram, min_length):
    """The fused per-row candidate kernel (forced tokens, log-softmax, EOS / n-gram bans, top-2k) drives the same beam search as
    the tensor-op implementation that tests/test_generate_oracle.py pins to the reference: identical token histories, beam
    re-ranking, hypothesis pools and done flags over a whole frame on random logits (low-entropy rows so that hypotheses
    finish and n-grams repeat)."""
    from multimodalsum_b200.generation import BeamSearch
    torch.manual_seed(20 + k)
    B, L = 6, 24
    N = B * k
    mk = lambda: BeamSearch(B, V, torch.device("cuda"), k, L, min_length, 1.0, ngram, True, 1, 0, 2)
    a, b = mk(), mk()
    b.use_kernel = False
    ld = (V + 3) // 4 * 4
    for step in range(L - 1):
        logits = torch.zeros(N, ld, device="cuda")[:, :V]
        logits.copy_(torch.randn(N, V, device="cuda") * 3)
        logits[:, 2] += 4.0                          # EOS is competitive: hypotheses finish along the way
        logits[:, 3:6] += 6.0                        # a few dominant tokens: repeated n-grams get banned
        ia = a.advance(logits)
        ib = b.advance(logits)
        assert a.kernel_path and not b.kernel_path
        assert torch.equal(a.done, b.done), step
        live = (~a.done).repeat_interleave(k)
        # (beam_idx itself may differ where candidates tie exactly - the k-1 dead beams of the first step - the histories not)
        assert ia.shape == ib.shape and torch.equal(a.ids[live], b.ids[live]), step
        assert torch.allclose(a.beam_scores[live], b.beam_scores[live], rtol=1e-5, atol=1e-4), step
        assert torch.equal(a.pool_n, b.pool_n) and torch.equal(a.pool_len, b.pool_len) and torch.equal(a.pool_tok, b.pool_tok), step
        fin = a.slot_ids[None, :] < a.pool_n[:, None]
        assert torch.allclose(a.pool_score[fin], b.pool_score[fin], rtol=1e-5, atol=1e-4), step
    assert torch.equal(a.finalize(), b.finalize())
