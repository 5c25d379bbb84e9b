"""Drop-in `nn.Module` surface of the reference for the training-step path.

Same class names, constructor meaning, forward signatures, return tuples and `state_dict` keys/shapes as
  MultimodalSum                         src/multimodal_train.py:111-193
  BartForMultiEncConditionalGeneration  src/transformer/modeling_multimodalsum.py:2181-2292 (parameter container here)
  YelpTableEncoder / AmazonTableEncoder src/table_encoder.py
  Resnet (projection head only)         src/img_encoder.py:26,39-40  (pooled stage-3 features in, trunk out of scope)
  LabelSmoothingLoss semantics          src/utils.py:22-38 (fused into the step)
The modules only HOLD parameters (fp32 masters, re-pointed into the engine's flat arena on first use); all compute
runs in the CUDA kernels behind `StepEngine`.  There is no eager/CPU fallback: without a CUDA device and the built
library, forward raises.
"""
import os

import torch
import torch.nn as nn

from .engine import StepEngine
from .synth import Batch, ModelConfig


def _init_linear(m, std):
    m.weight.data.normal_(mean=0.0, std=std)
    if m.bias is not None:
        m.bias.data.zero_()


class _Attention(nn.Module):
    def __init__(self, d, cross=False, multimodal=False):
        super().__init__()
        # registration order k, v, q, out as in the reference (modeling_multimodalsum.py:695-698)
        self.k_proj = nn.Linear(d, d)
        self.v_proj = nn.Linear(d, d)
        self.q_proj = nn.Linear(d, d)
        self.out_proj = nn.Linear(d, d)
        if cross and multimodal:
            self.alpha_proj = nn.Linear(2 * d, d)
            self.beta_proj = nn.Linear(2 * d, d)


class _EncoderLayer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        d = cfg.d_model
        self.self_attn = _Attention(d)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, cfg.ffn_dim)
        self.fc2 = nn.Linear(cfg.ffn_dim, d)
        self.final_layer_norm = nn.LayerNorm(d)


class _DecoderLayer(nn.Module):
    def __init__(self, cfg, multimodal):
        super().__init__()
        d = cfg.d_model
        self.self_attn = _Attention(d)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.encoder_attn = _Attention(d, cross=True, multimodal=multimodal)
        self.encoder_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, cfg.ffn_dim)
        self.fc2 = nn.Linear(cfg.ffn_dim, d)
        self.final_layer_norm = nn.LayerNorm(d)


class BartEncoder(nn.Module):
    def __init__(self, cfg, embed_tokens):
        super().__init__()
        self.embed_tokens = embed_tokens
        self.embed_positions = nn.Embedding(cfg.max_position_embeddings + 2, cfg.d_model, padding_idx=cfg.pad_token_id)
        self.layers = nn.ModuleList([_EncoderLayer(cfg) for _ in range(cfg.encoder_layers)])
        self.layernorm_embedding = nn.LayerNorm(cfg.d_model)


class BartDecoder(nn.Module):
    def __init__(self, cfg, embed_tokens, multimodal):
        super().__init__()
        self.embed_tokens = embed_tokens
        self.rating_embeddings = nn.Parameter(torch.empty(cfg.d_model).normal_(mean=0.0, std=cfg.init_std))
        self.embed_positions = nn.Embedding(cfg.max_position_embeddings + 2, cfg.d_model, padding_idx=cfg.pad_token_id)
        self.layers = nn.ModuleList([_DecoderLayer(cfg, multimodal) for _ in range(cfg.decoder_layers)])
        self.layernorm_embedding = nn.LayerNorm(cfg.d_model)


class BartModel(nn.Module):
    def __init__(self, cfg, multimodal):
        super().__init__()
        self.shared = nn.Embedding(cfg.vocab_size, cfg.d_model, padding_idx=cfg.pad_token_id)
        self.encoder = BartEncoder(cfg, self.shared)
        self.decoder = BartDecoder(cfg, self.shared, multimodal)


class BartForMultiEncConditionalGeneration(nn.Module):
    """Parameter container with the reference's keys; `multimodal=False` gives BartForEncConditionalGeneration's layout."""

    def __init__(self, config: ModelConfig, multimodal=True):
        super().__init__()
        self.config = config
        self.model = BartModel(config, multimodal)
        self.register_buffer("final_logits_bias", torch.zeros((1, config.vocab_size)))
        self.apply(self._init_weights)

    def _init_weights(self, m):
        # PretrainedBartModel._init_weights, modeling_multimodalsum.py:188-199
        std = self.config.init_std
        if isinstance(m, nn.Linear):
            _init_linear(m, std)
        elif isinstance(m, nn.Embedding):
            m.weight.data.normal_(mean=0.0, std=std)
            if m.padding_idx is not None:
                m.weight.data[m.padding_idx].zero_()


BartForEncConditionalGeneration = lambda config: BartForMultiEncConditionalGeneration(config, multimodal=False)  # noqa: E731


class YelpTableEncoder(nn.Module):
    def __init__(self, bart_embedding):
        super().__init__()
        self.bart_embedding = bart_embedding
        self.rating_embedding = nn.Linear(4, 1024, bias=False)
        self.hours_embedding = nn.Linear(4, 1024, bias=False)
        self.fc = nn.Linear(2048, 1024)
        self.linear = nn.Linear(1024, 1024, bias=False)


class AmazonTableEncoder(nn.Module):
    def __init__(self, bart_embedding):
        super().__init__()
        self.bart_embedding = bart_embedding
        self.price_embedding = nn.Linear(11, 1024, bias=False)
        self.rating_embedding = nn.Linear(4, 1024, bias=False)
        self.fc = nn.Linear(2048, 1024)
        self.linear = nn.Linear(1024, 1024, bias=False)


class Resnet(nn.Module):
    """Projection head of src/img_encoder.py (`linear`, 1024 -> d_model, no bias).  The ResNet-101 trunk is outside the
    hot path (north star): `img` inputs are its pooled stage-3 features [B, max_imgs, 196, 1024]."""

    def __init__(self, embedding_dim):
        super().__init__()
        self.linear = nn.Linear(1024, embedding_dim, bias=False)


class _StepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, module, batch, label_smoothing):
        eng = module.engine
        loss = eng.forward(batch, label_smoothing, training=module.training)
        ctx.engine = eng
        return loss.reshape(()).clone()

    @staticmethod
    def backward(ctx, grad_out):
        ctx.engine.backward(grad_out.contiguous().float())
        return None, None, None, None


class MultimodalSum(nn.Module):
    """`MultimodalSum(bart_pretrained, table_pretrained, img_pretrained, TableEncoder)` as in src/multimodal_train.py:111-122.

    `*_pretrained` are local checkpoint directories holding `pytorch_model.bin` (no hub download in this build; `None`
    = random init by the reference's recipe).  `config` selects the model dims (default: cfg/bart-large.json values).
    `label_smoothing` replaces the reference's module-global `args.label_smoothing` (default 0.1, :433)."""

    def __init__(self, bart_pretrained=None, table_pretrained=None, img_pretrained=None, TableEncoder=YelpTableEncoder,
                 config: ModelConfig = None, label_smoothing=0.1):
        super().__init__()
        ds = "yelp" if TableEncoder is YelpTableEncoder else "amazon"
        cfg = config if config is not None else ModelConfig(dataset=ds)
        if cfg.dataset != ds:
            raise ValueError("config.dataset %r does not match the TableEncoder (%s)" % (cfg.dataset, ds))
        self.cfg = cfg
        self.label_smoothing = label_smoothing
        self.bart_model = BartForMultiEncConditionalGeneration(cfg)
        if bart_pretrained is not None:
            sd = torch.load(os.path.join(bart_pretrained, "pytorch_model.bin"), map_location="cpu")
            self.bart_model.load_state_dict(sd, strict=False)  # alpha/beta_proj are authorized missing keys (:2183)
        self.table_encoder = TableEncoder(self.bart_model.model.shared)
        if table_pretrained is not None:
            self.table_encoder.load_state_dict(torch.load(os.path.join(table_pretrained, "pytorch_model.bin"), map_location="cpu"))
        self.img_encoder = Resnet(cfg.d_model)
        if img_pretrained is not None:
            self.img_encoder.load_state_dict(torch.load(os.path.join(img_pretrained, "pytorch_model.bin"), map_location="cpu"),
                                             strict=False)
        self.engine = None

    # -- engine plumbing ------------------------------------------------------------------------------
    def _ensure_engine(self, device):
        if self.engine is None:
            if device.type != "cuda":
                raise RuntimeError("mmsum_b200 has no CPU path: move the module and its inputs to a CUDA device")
            eng = StepEngine(self.cfg, device)
            eng.bind(self.named_parameters())
            eng.final_logits_bias = self.bart_model.final_logits_bias
            object.__setattr__(self, "engine", eng)
        return self.engine

    def forward(self, reviews, reviews_mask, reviews_rating, field, field_value, img, img_mask, **unused):
        """-> (loss,)  — same arguments as the reference (src/multimodal_train.py:124-139); `img` holds pooled
        ResNet-101 stage-3 features [B, max_imgs, 196, 1024] (fp32 or bf16)."""
        eng = self._ensure_engine(reviews.device)
        batch = Batch(reviews, reviews_mask, reviews_rating, field, list(field_value), img, img_mask)
        eng._batch = batch
        loss = _StepFn.apply(eng.anchor, self, batch, self.label_smoothing)
        return (loss,)


class TextSupervised(nn.Module):
    """Text-only step of src/text_pretrain.py:66-113 (BASELINE config 1): same engine, one text memory,
    plain cross-entropy unless `label_smoothing` is given."""

    def __init__(self, bart_pretrained=None, config: ModelConfig = None, label_smoothing=None):
        super().__init__()
        cfg = config if config is not None else ModelConfig(dataset="text")
        if cfg.dataset != "text":
            raise ValueError("TextSupervised needs config.dataset == 'text'")
        self.cfg = cfg
        self.label_smoothing = label_smoothing
        self.bart_model = BartForMultiEncConditionalGeneration(cfg, multimodal=False)
        if bart_pretrained is not None:
            self.bart_model.load_state_dict(torch.load(os.path.join(bart_pretrained, "pytorch_model.bin"), map_location="cpu"),
                                            strict=False)
        self.engine = None

    _ensure_engine = MultimodalSum._ensure_engine

    def forward(self, reviews, reviews_mask, reviews_rating, **unused):
        eng = self._ensure_engine(reviews.device)
        batch = Batch(reviews, reviews_mask, reviews_rating)
        eng._batch = batch
        loss = _StepFn.apply(eng.anchor, self, batch, self.label_smoothing)
        return (loss,)
