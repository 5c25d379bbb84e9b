"""The reference's module surface on the GPU: every sub-module forward of the drop-in against the oracle
function of the same name, and the body of src/test.py:152-158 run verbatim against the drop-in.

Tolerances: integer outputs (masks, token ids) exact; bf16 activations within 2e-2 of the fp32 oracle relative to the
tensor's max magnitude; logits within 3e-2 (bf16 through 2 + 2 layers); losses 2e-3 relative."""
import json
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN_DIR, load_golden
from multimodalsum_b200.synth import ModelConfig, make_batch, make_state_dict

pytestmark = pytest.mark.gpu


def _close(a, b, rtol, name=""):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item()
    scale = max(b.abs().max().item(), 1e-6)
    assert err <= rtol * scale, "%s: max err %.4g vs scale %.4g" % (name, err, scale)


def _model(cfg, sd, cls=None, **kw):
    from multimodalsum_b200.modules import AmazonTableEncoder, MultimodalSum, YelpTableEncoder
    if cls is None:
        model = MultimodalSum(TableEncoder=YelpTableEncoder if cfg.dataset == "yelp" else AmazonTableEncoder, config=cfg)
    else:
        model = cls(config=cfg, **kw)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model.cuda().eval()


def _oracle_params(sd):
    torch.backends.cuda.matmul.allow_tf32 = False
    return {k: v.cuda() for k, v in sd.items()}


@pytest.mark.parametrize("S", [128, 150])
def test_bart_encoder_forward(S):
    from oracle import mmsum_oracle as OR
    gold = load_golden("small_yelp")
    cfg, sd = gold["cfg"], gold["sd"]
    cfg.max_position_embeddings = 128
    if S > 128:
        cfg = ModelConfig(**dict(json.loads(str(np.load(os.path.join(GOLDEN_DIR, "gen_small_yelp_s150.npz"))["case"]))["cfg"]))
        sd = make_state_dict(cfg, seed=21)
    model = _model(cfg, sd)
    b = make_batch(cfg, 2, seed=5, n_reviews=3, max_imgs=1, seq_len=S, len_range=(S // 2, S - 10)).to("cuda")
    ids, mask = b.reviews.view(-1, S), b.reviews_mask.view(-1, S)
    out = model.bart_model.model.encoder(input_ids=ids, attention_mask=mask)
    assert isinstance(out, tuple) and out[0].shape == (6, S, cfg.d_model)
    ref = OR.encoder(_oracle_params(sd), cfg, ids, mask)
    valid = mask.bool()
    _close(out[0][valid], ref[valid], 2e-2, "encoder")        # pad rows are never consumed (masked as keys everywhere)


@pytest.mark.parametrize("name", ["small_yelp", "small_amazon"])
def test_table_encoder_and_image_head_forward(name):
    from oracle import mmsum_oracle as OR
    gold = load_golden(name)
    cfg, sd = gold["cfg"], gold["sd"]
    model = _model(cfg, sd)
    b = gold["batch"].to("cuda")
    p = _oracle_params(sd)
    emb, mask = model.table_encoder(b.field, b.field_value)
    ref_emb, ref_mask = (OR.yelp_table_encoder if cfg.dataset == "yelp" else OR.amazon_table_encoder)(p, b.field, b.field_value)
    assert mask.dtype == torch.bool and torch.equal(mask, ref_mask)
    _close(emb, ref_emb, 2e-2, "table")
    img = model.img_encoder(b.img)
    assert img.shape == b.img.shape[:3] + (cfg.d_model,)
    _close(img, torch.nn.functional.linear(b.img, p["img_encoder.linear.weight"]), 2e-2, "img")


@pytest.mark.parametrize("smoothing,V", [(0.1, 50265), (0.0, 777)])
def test_label_smoothing_loss_module(smoothing, V):
    from multimodalsum_b200.modules import LabelSmoothingLoss
    torch.manual_seed(0)
    rows = 300
    pred = (torch.randn(rows, V, device="cuda") * 2).to(torch.bfloat16).float().requires_grad_(True)   # bf16-representable logits
    target = torch.randint(0, V, (rows,), device="cuda")
    loss = LabelSmoothingLoss(V, smoothing=smoothing)(pred, target)
    loss.backward()
    g = pred.grad.clone()
    ref_in = pred.detach().clone().requires_grad_(True)
    logp = ref_in.log_softmax(-1)
    dist = torch.full_like(logp, smoothing / (V - 1)).scatter_(1, target[:, None], 1.0 - smoothing)
    ref = (-dist * logp).sum(-1).mean()
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    assert (g - ref_in.grad).abs().max().item() <= 1e-2 * ref_in.grad.abs().max().item() + 1e-9    # bf16 gradient storage


def test_multienc_forward_logits_and_get_multimodal_outputs():
    """get_multimodal_outputs 7-tuple -> BartForMultiEncConditionalGeneration.forward(labels=...) -> logits, vs the oracle's
    multimodal_memories / decoder / lm_logits on the same inputs (one leave-one-out pass written as the reference does)."""
    from oracle import mmsum_oracle as OR
    gold = load_golden("small_yelp_gates_open")
    cfg, sd = gold["cfg"], gold["sd"]
    model = _model(cfg, sd)
    b = gold["batch"].to("cuda")
    outs = model.get_multimodal_outputs(b.reviews, b.reviews_mask, b.field, b.field_value, b.img, b.img_mask)
    n_reviews, th, tm, tabh, tabm, ih, im = outs
    B, R, S = b.reviews.shape
    assert n_reviews == R and th.shape == (B, R, S, cfg.d_model) and tabh.shape == (B, 1, 47, cfg.d_model)
    assert tabm.shape == (B, 1, 47) and ih.shape == (B, b.img.shape[1], 196, cfg.d_model) and im.shape == ih.shape[:3]
    p = _oracle_params(sd)
    text, text_valid, table, table_valid, img, img_valid = OR.multimodal_memories(p, cfg, b)
    assert torch.equal(tabm, table_valid) and torch.equal(im, img_valid)
    _close(tabh, table, 2e-2, "table memory")
    _close(ih, img, 2e-2, "image memory")
    others = [0, 2]
    rating_diff = (b.reviews_rating[:, 1] - b.reviews_rating[:, others].mean(dim=1)).unsqueeze(1)
    labels = b.reviews[:, 1, :]
    logits = model.bart_model(th[:, others], tm[:, others], tabh, tabm, ih, im, rating_diff=rating_diff, labels=labels)[0]
    assert logits.shape == (B, S, cfg.vocab_size) and logits.dtype == torch.float32
    dec_ids = OR.shift_tokens_right(labels, cfg.pad_token_id, cfg.bos_token_id, cfg.eos_token_id)
    x = OR.decoder(p, cfg, dec_ids, [text[:, others], table, img], [text_valid[:, others], table_valid, img_valid], rating_diff)
    ref = OR.lm_logits(x, p)
    _close(logits, ref, 3e-2, "lm logits")
    lp, lr = torch.log_softmax(logits, -1), torch.log_softmax(ref, -1)
    assert (lp - lr).abs().max().item() <= 0.05


def test_text_only_enc_forward_logits():
    from multimodalsum_b200.modules import TextSupervised
    from oracle import mmsum_oracle as OR
    gold = load_golden("small_text")
    cfg, sd = gold["cfg"], gold["sd"]
    model = _model(cfg, sd, TextSupervised)
    b = gold["batch"].to("cuda")
    B, R, S = b.reviews.shape
    th = model.bart_model.model.encoder(input_ids=b.reviews.view(B * R, S), attention_mask=b.reviews_mask.view(B * R, S))[0]
    th = th.reshape(B, R, S, -1)
    labels = b.reviews[:, 0]
    rd = torch.zeros(B, 1, device="cuda")
    logits = model.bart_model(th[:, 1:], rd, b.reviews_mask[:, 1:], labels=labels)[0]
    p = _oracle_params(sd)
    text = OR.encoder(p, cfg, b.reviews.view(B * R, S), b.reviews_mask.view(B * R, S)).view(B, R, S, -1)
    dec_ids = OR.shift_tokens_right(labels, cfg.pad_token_id, cfg.bos_token_id, cfg.eos_token_id)
    x = OR.decoder(p, cfg, dec_ids, text[:, 1:], b.reviews_mask[:, 1:].bool(), rd)
    _close(logits, OR.lm_logits(x, p), 3e-2, "text-only logits")


def test_reference_test_py_body_runs_verbatim_against_the_drop_in():
    """src/test.py:152-158, statement for statement, with `model` = the drop-in MultimodalSum: the generated ids must equal
    what the UNMODIFIED reference produced on the same weights / inputs (tie-free golden)."""
    z = np.load(os.path.join(GOLDEN_DIR, "gen_small_yelp_biased.npz"), allow_pickle=False)
    case = json.loads(str(z["case"]))
    cfg = ModelConfig(**case["cfg"])
    sd = make_state_dict(cfg, **case["sd"])
    bk = dict(case["batch"])
    batch = make_batch(cfg, bk.pop("B"), **bk).to("cuda")
    model = _model(cfg, sd)
    reviews, reviews_mask, field, field_value, img, img_mask = (batch.reviews, batch.reviews_mask, batch.field, batch.field_value,
                                                                batch.img, batch.img_mask)
    args = type("Args", (), dict(num_beams=case["gen"]["num_beams"], length_penalty=case["gen"]["length_penalty"],
                                 max_length=case["gen"]["max_length"]))
    # ---- verbatim from src/test.py:152-158 -------------------------------------------------------------------------
    with torch.no_grad():
        _, text_hiddens, text_attention_mask, table_hiddens, table_attention_mask, img_hiddens, img_attention_mask = \
        model.get_multimodal_outputs(reviews, reviews_mask, field, field_value, img, img_mask)
        rating_diff = torch.zeros([text_hiddens.size(0), 1], device=text_hiddens.device)
        generated = model.bart_model.generate(text_hiddens, text_attention_mask, table_hiddens, table_attention_mask, img_hiddens, img_attention_mask, 
                                              rating_diff=rating_diff, num_beams=args.num_beams, length_penalty=args.length_penalty, max_length=args.max_length,
                                              no_repeat_ngram_size=3, early_stopping=True)
    # ------------------------------------------------------------------------------------------------------------------
    assert torch.equal(generated.cpu(), torch.from_numpy(z["tokens"])), (generated.cpu().tolist(), z["tokens"].tolist())


def test_standalone_bart_container_owns_its_engine():
    """BartForMultiEncConditionalGeneration built on its own (no table / image encoder modules) runs forward from given memories."""
    from multimodalsum_b200.modules import BartForMultiEncConditionalGeneration
    gold = load_golden("small_yelp")
    cfg, sd = gold["cfg"], gold["sd"]
    full = _model(cfg, sd)
    b = gold["batch"].to("cuda")
    _, th, tm, tabh, tabm, ih, im = full.get_multimodal_outputs(b.reviews, b.reviews_mask, b.field, b.field_value, b.img, b.img_mask)
    labels = b.reviews[:, 0]
    ref = full.bart_model(th[:, 1:], tm[:, 1:], tabh, tabm, ih, im, labels=labels)[0]
    alone = BartForMultiEncConditionalGeneration(cfg)
    alone.load_state_dict({k[len("bart_model."):]: v for k, v in sd.items() if k.startswith("bart_model.")})
    alone = alone.cuda().eval()
    out = alone(th[:, 1:], tm[:, 1:], tabh, tabm, ih, im, labels=labels)[0]
    assert torch.equal(out, ref)


def test_submodule_without_a_model_fails_loudly():
    from multimodalsum_b200.modules import YelpTableEncoder
    enc = YelpTableEncoder(torch.nn.Embedding(10, 1024)).cuda()
    with pytest.raises(RuntimeError):
        enc(torch.zeros(47, 6, dtype=torch.long, device="cuda"), [])
