"""Deterministic synthetic weights and Yelp/Amazon-shaped inputs (no network, no reference needed).

Weights follow the reference's random-init recipe distributionally (N(0, init_std) for Linear/Embedding
weights, zero pad row; src/transformer/modeling_multimodalsum.py:188-199, :514-515) but every tensor is drawn
from its own `torch.Generator` seeded by crc32(name), so the same state_dict can be rebuilt anywhere (this
container, the GPU box, the golden-fixture script) without replaying the reference's constructor RNG order.
`perturb=True` additionally gives biases / LayerNorm affine parameters non-trivial values so their gradient
paths are exercised by the parity tests.

Inputs follow SURVEY.md §8(d) / App. G: token frames of 128 with EOS at len-1 and PAD(1) after, ratings 1..5,
Yelp table (47 fields) or Amazon table (133 fields), post-ReLU image features [B, max_imgs, 196, 1024].
"""
import zlib
from dataclasses import dataclass, field as dc_field

import torch


@dataclass
class ModelConfig:
    """The subset of cfg/bart-large.json the hot path reads."""
    d_model: int = 1024
    encoder_layers: int = 12
    decoder_layers: int = 12
    heads: int = 16
    ffn_dim: int = 4096
    vocab_size: int = 50265
    max_position_embeddings: int = 1024
    pad_token_id: int = 1
    bos_token_id: int = 0
    eos_token_id: int = 2
    dropout: float = 0.1
    init_std: float = 0.02
    # 'yelp' | 'amazon': multimodal step (text + table + image memory, gated fusion)      src/multimodal_train.py
    # 'text': text-only step (config 1)                                                   src/text_pretrain.py
    # 'img' | 'table_yelp' | 'table_amazon': single-modality pretraining stages            src/img_pretrain.py, src/table_pretrain.py
    dataset: str = "yelp"

    @property
    def head_dim(self):
        return self.d_model // self.heads

    @property
    def multimodal(self):
        """Decoder cross-attention fuses three modalities with alpha/beta gates (BartForMultiEncConditionalGeneration)."""
        return self.dataset in ("yelp", "amazon")

    @property
    def text_memory(self):
        """The decoder attends to encoded reviews (leave-one-out); false for the img / table pretraining stages."""
        return self.dataset in ("yelp", "amazon", "text")

    @property
    def table(self):
        return {"yelp": "yelp", "amazon": "amazon", "table_yelp": "yelp", "table_amazon": "amazon"}.get(self.dataset)

    @property
    def image(self):
        return self.dataset in ("yelp", "amazon", "img")

    def to_reference_dict(self):
        return dict(d_model=self.d_model, encoder_layers=self.encoder_layers, decoder_layers=self.decoder_layers,
                    encoder_attention_heads=self.heads, decoder_attention_heads=self.heads,
                    encoder_ffn_dim=self.ffn_dim, decoder_ffn_dim=self.ffn_dim, vocab_size=self.vocab_size,
                    max_position_embeddings=self.max_position_embeddings, dropout=self.dropout)


BART_LARGE = ModelConfig()


def _draw(name, shape, std, seed, mean=0.0):
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return torch.randn(shape, generator=g, dtype=torch.float32) * std + mean


def param_shapes(cfg: ModelConfig):
    """Ordered {name: shape} of the reference state_dict for this path."""
    D, F, V = cfg.d_model, cfg.ffn_dim, cfg.vocab_size
    s = {}
    bm = "bart_model."
    s[bm + "final_logits_bias"] = (1, V)
    s[bm + "model.shared.weight"] = (V, D)

    def attn(prefix, cross):
        for p in ("k_proj", "v_proj", "q_proj", "out_proj"):
            s[prefix + p + ".weight"] = (D, D)
            s[prefix + p + ".bias"] = (D,)
        if cross and cfg.multimodal:
            for p in ("alpha_proj", "beta_proj"):
                s[prefix + p + ".weight"] = (D, 2 * D)
                s[prefix + p + ".bias"] = (D,)

    def ln(prefix):
        s[prefix + ".weight"] = (D,)
        s[prefix + ".bias"] = (D,)

    for side, nl in (("encoder", cfg.encoder_layers), ("decoder", cfg.decoder_layers)):
        pre = bm + "model.%s." % side
        s[pre + "embed_tokens.weight"] = (V, D)  # alias of shared
        if side == "decoder":
            s[pre + "rating_embeddings"] = (D,)
        s[pre + "embed_positions.weight"] = (cfg.max_position_embeddings + 2, D)
        for i in range(nl):
            lp = pre + "layers.%d." % i
            attn(lp + "self_attn.", False)
            ln(lp + "self_attn_layer_norm")
            if side == "decoder":
                attn(lp + "encoder_attn.", True)
                ln(lp + "encoder_attn_layer_norm")
            s[lp + "fc1.weight"] = (F, D)
            s[lp + "fc1.bias"] = (F,)
            s[lp + "fc2.weight"] = (D, F)
            s[lp + "fc2.bias"] = (D,)
            ln(lp + "final_layer_norm")
        ln(pre + "layernorm_embedding")
    if cfg.table is not None:
        t = "table_encoder."
        s[t + "bart_embedding.weight"] = (V, D)  # alias of shared
        if cfg.table == "yelp":
            s[t + "rating_embedding.weight"] = (D, 4)
            s[t + "hours_embedding.weight"] = (D, 4)
        else:
            s[t + "price_embedding.weight"] = (D, 11)
            s[t + "rating_embedding.weight"] = (D, 4)
        s[t + "fc.weight"] = (D, 2 * D)
        s[t + "fc.bias"] = (D,)
        s[t + "linear.weight"] = (D, D)
    if cfg.image:
        s["img_encoder.linear.weight"] = (D, 1024)
    return s


ALIASES = ("model.encoder.embed_tokens.weight", "model.decoder.embed_tokens.weight", "table_encoder.bart_embedding.weight")


def make_state_dict(cfg: ModelConfig, seed=0, perturb=True, gates_open=False, logits_bias_std=0.0):
    """fp32 CPU state_dict with the reference's keys.  `gates_open` = stress init of SURVEY App. G (alpha/beta
    bias +1, weights x5) so table / image branches contribute O(1) to the loss."""
    sd = {}
    shared = None
    for name, shape in param_shapes(cfg).items():
        if name.endswith(ALIASES):
            sd[name] = shared
            continue
        if name.endswith("final_logits_bias"):
            sd[name] = _draw(name, shape, logits_bias_std, 77) if logits_bias_std > 0 else torch.zeros(shape)
        elif "layer_norm" in name or "layernorm_embedding" in name:
            if name.endswith(".weight"):
                sd[name] = _draw(name, shape, 0.05 if perturb else 0.0, seed, mean=1.0)
            else:
                sd[name] = _draw(name, shape, 0.05 if perturb else 0.0, seed)
        elif name.endswith(".bias"):
            sd[name] = _draw(name, shape, cfg.init_std if perturb else 0.0, seed)
        else:
            sd[name] = _draw(name, shape, cfg.init_std, seed)
        if name.endswith("model.shared.weight"):
            sd[name][cfg.pad_token_id].zero_()
            shared = sd[name]
        if name.endswith("embed_positions.weight"):
            sd[name][cfg.pad_token_id].zero_()
        if gates_open and ("alpha_proj" in name or "beta_proj" in name):
            sd[name] = sd[name] * 5.0 + (1.0 if name.endswith(".bias") else 0.0)
    return sd


def make_grads(cfg: ModelConfig, seed=0, std=1e-3):
    """Deterministic synthetic gradients keyed like `named_parameters()` (optimizer parity tests: the same gradients can be
    rebuilt in the build container for the reference's AdamW and on the GPU box for the fused kernel)."""
    return {name: _draw(name + "#grad", shape, std, seed) for name, shape in param_shapes(cfg).items()
            if not name.endswith(ALIASES) and not name.endswith("final_logits_bias")}


@dataclass
class Batch:
    reviews: torch.Tensor          # i64 [B, R, S]
    reviews_mask: torch.Tensor     # i64 [B, R, S]  (1 = valid)
    reviews_rating: torch.Tensor   # f32 [B, R]
    field: torch.Tensor = None     # i64 [47,6] yelp / [6,1] amazon
    field_value: list = dc_field(default_factory=list)
    img: torch.Tensor = None       # f32 [B, max_imgs, 196, 1024] pooled ResNet-101 stage-3 features
    img_mask: torch.Tensor = None  # bool [B, max_imgs]
    labels: torch.Tensor = None    # i64 [B, S]: decoder targets of the img / table pretraining stages (reviews is None there)
    max_review_len: int = None     # optional host-side hint: no review has a valid token at position >= max_review_len (the
                                   # engine trims the encoder frames to it; without the hint it reads the mask back once)

    def to(self, device, non_blocking=False):
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return Batch(mv(self.reviews), mv(self.reviews_mask), mv(self.reviews_rating), mv(self.field),
                     [mv(t) for t in self.field_value], mv(self.img), mv(self.img_mask), mv(self.labels), self.max_review_len)

    def pin(self):
        pv = lambda t: None if t is None else t.pin_memory()
        return Batch(pv(self.reviews), pv(self.reviews_mask), pv(self.reviews_rating), pv(self.field),
                     [pv(t) for t in self.field_value], pv(self.img), pv(self.img_mask), pv(self.labels), self.max_review_len)

    def with_length_hint(self):
        """Fill `max_review_len` from the mask (a host-side pass when the batch still lives on the CPU: what a collate function
        knows anyway; a device read-back otherwise)."""
        if self.reviews_mask is not None:
            m = self.reviews_mask.ne(0)
            S = m.shape[-1]
            self.max_review_len = int((m.to(torch.int32) * torch.arange(1, S + 1, device=m.device, dtype=torch.int32)).max().item())
        return self

    def tensors(self):
        return [t for t in [self.reviews, self.reviews_mask, self.reviews_rating, self.field, self.img, self.img_mask, self.labels]
                + list(self.field_value) if t is not None]

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors())


def _pad_tail(tok, g, lo=1):
    """Give every last-dim token row a random valid length in [lo, L] followed by PAD(1)."""
    L = tok.shape[-1]
    lens = torch.randint(lo, L + 1, tok.shape[:-1], generator=g).unsqueeze(-1)
    return torch.where(torch.arange(L) < lens, tok, torch.ones_like(tok))


def make_batch(cfg: ModelConfig, B, seed=1, n_reviews=9, seq_len=128, fixed_len=None, max_imgs=None,
               n_valid_imgs=None, len_range=None):
    """Synthetic batch of B businesses.  fixed_len: all reviews have that many valid tokens (throughput runs:
    100 yelp / 70 amazon); otherwise U{60..100} (yelp) / U{45..70} (amazon), clipped to seq_len."""
    g = torch.Generator().manual_seed(seed)
    V = cfg.vocab_size
    R, S = n_reviews, seq_len
    rev = torch.randint(3, V, (B, R, S), generator=g)
    if fixed_len is not None:
        lens = torch.full((B, R), min(fixed_len, S), dtype=torch.long)
    else:
        lo, hi = len_range if len_range is not None else ((45, 70) if cfg.dataset == "amazon" else (60, 100))
        lo, hi = min(lo, S), min(hi, S)
        lens = torch.randint(lo, hi + 1, (B, R), generator=g)
    pos = torch.arange(S)
    rev = torch.where(pos < lens.unsqueeze(-1), rev, torch.ones_like(rev))
    rev = torch.where(pos == (lens - 1).unsqueeze(-1), torch.full_like(rev, cfg.eos_token_id), rev)
    mask = (rev != cfg.pad_token_id).long()
    rating = torch.randint(1, 6, (B, R), generator=g).float()
    batch = Batch(rev, mask, rating)
    if cfg.dataset == "text":
        return batch
    if not cfg.text_memory:
        # pretraining stages: one target review per business, memory = its images / its table only
        batch = Batch(None, None, None, labels=rev[:, 0].contiguous())
    if cfg.table == "yelp":
        batch.field = _pad_tail(torch.randint(3, V, (47, 6), generator=g), g)
        name = _pad_tail(torch.randint(3, V, (B, 24), generator=g), g)
        category = _pad_tail(torch.randint(3, V, (B, 6, 12), generator=g), g, lo=0)
        str_cat = _pad_tail(torch.randint(3, V, (B, 5, 3), generator=g), g, lo=0)
        str_bool = _pad_tail(torch.randint(3, V, (B, 32, 1), generator=g), g, lo=0)
        rbits = torch.randint(0, 2, (B, 4), generator=g)
        hsel = torch.randint(0, 5, (B, 7), generator=g)  # 4 = closed / missing -> all-zero row
        hours = torch.zeros(B, 7, 4, dtype=torch.long)
        hours.scatter_(2, hsel.clamp(max=3).unsqueeze(-1), (hsel < 4).long().unsqueeze(-1))
        batch.field_value = [name, category, str_cat, str_bool, rbits, hours]
        mi = 10 if max_imgs is None else max_imgs
    elif cfg.table == "amazon":
        batch.field = torch.randint(3, V, (6, 1), generator=g)
        price = torch.randint(0, 2, (B, 11), generator=g)
        rbits = torch.randint(0, 2, (B, 4), generator=g)
        brand = _pad_tail(torch.randint(3, V, (B, 12), generator=g), g, lo=0)
        name = _pad_tail(torch.randint(3, V, (B, 32), generator=g), g, lo=0)
        category = _pad_tail(torch.randint(3, V, (B, 3, 8, 12), generator=g), g, lo=0)
        desc = _pad_tail(torch.randint(3, V, (B, 128), generator=g), g, lo=0)
        batch.field_value = [price, rbits, brand, name, category, desc]
        mi = 1 if max_imgs is None else max_imgs
    else:
        mi = 10 if max_imgs is None else max_imgs
    if not cfg.image:
        return batch
    batch.img = torch.relu(torch.randn(B, mi, 196, 1024, generator=g))
    if n_valid_imgs is None:
        k = torch.randint(0, mi + 1, (B,), generator=g)
    else:
        k = torch.full((B,), n_valid_imgs, dtype=torch.long)
    batch.img_mask = torch.arange(mi).unsqueeze(0) < k.unsqueeze(1)
    return batch


class SyntheticDataset(torch.utils.data.Dataset):
    """Map-style dataset with the item layout of the reference's `MultimodalDataset.__getitem__` (src/multimodal_train.py:63-86):
    one business per item — (reviews [R,S], reviews_mask, reviews_rating [R], <six table fields>, img [max_imgs,196,1024],
    img_mask [max_imgs]) — so that `DataLoader(..., pin_memory=True)` collates the 11-tuple the prefetchers expect, and `.field`
    [47,6] / [6,1] as `data_train.field` (:59-62).  Items are generated from `seed + idx` on demand (deterministic, no storage);
    image entries are pooled ResNet-101 stage-3 features, not pixels (the trunk is outside this path)."""

    def __init__(self, cfg: ModelConfig, n_businesses, seed=0, **batch_kw):
        if cfg.dataset not in ("yelp", "amazon"):
            raise ValueError("SyntheticDataset mirrors the multimodal stage: cfg.dataset must be 'yelp' or 'amazon'")
        self.cfg, self.n, self.seed, self.kw = cfg, int(n_businesses), int(seed), batch_kw
        self.field = make_batch(cfg, 1, seed=seed, **batch_kw).field
        self.epoch = 0

    def __len__(self):
        return self.n

    def set_epoch(self):
        """The reference re-draws its leave-one-out sampling between epochs (src/train_utils.py:72-73); here: new synthetic items."""
        self.epoch += 1

    def __getitem__(self, idx):
        if not 0 <= idx < self.n:
            raise IndexError(idx)
        b = make_batch(self.cfg, 1, seed=self.seed + 1 + idx + self.epoch * self.n, **self.kw)
        return (b.reviews[0], b.reviews_mask[0], b.reviews_rating[0], *[v[0] for v in b.field_value], b.img[0], b.img_mask[0])
