"""Beam-search generation (BASELINE config 5) on the CUDA kernels of the training step.

Reference: BartForMultiEncConditionalGeneration.generate / _generate_beam_search / adjust_logits_during_generation
(src/transformer/modeling_multimodalsum.py:2295-3101), postprocess_next_token_scores / calc_banned_ngram_tokens /
BeamHypotheses (src/transformer/generation_utils.py:57-99, 848-868, 948-993), driven by src/test.py:152-158.

Round-1 design (correct first, cached decode next):
  * the multimodal memory is encoded ONCE and its cross-attention K|V is projected ONCE per decoder layer and kept
    UN-EXPANDED per business — all beams of a business attend to the same memory (the reference expands every memory
    num_beams times and re-gathers the expanded K/V cache with index_select at every token, :2598-2627, :3004-3010);
  * every decode step re-runs the decoder over the current prefix inside a 128-position causal frame and reads the
    logits of the last position.  This is mathematically identical to the reference's cached single-token step (causal
    self-attention over the same prefix, position = cur_len-1) and reuses the training kernels unchanged; the
    self-attention K/V cache (one-token steps) is the round-2 item;
  * beam bookkeeping (n-gram blocking, hypothesis heaps, early stopping) stays host-side Python exactly as in the
    reference; log-softmax / top-k over [B*beams, V] use torch on the device.
Review frames up to 208 tokens are supported (src/test.py uses 158): encoder frames are padded to a multiple of 128 and
handled as two query tiles.
"""
import math

import torch

from . import ops


class BeamHypotheses:
    """generation_utils.py:948-993."""

    def __init__(self, num_beams, max_length, length_penalty, early_stopping):
        self.max_length = max_length - 1
        self.length_penalty = length_penalty
        self.early_stopping = early_stopping
        self.num_beams = num_beams
        self.beams = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / len(hyp) ** self.length_penalty
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp))
            if len(self) > self.num_beams:
                sorted_scores = sorted([(s, idx) for idx, (s, _) in enumerate(self.beams)])
                del self.beams[sorted_scores[0][1]]
                self.worst_score = sorted_scores[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self) < self.num_beams:
            return False
        if self.early_stopping:
            return True
        return self.worst_score >= best_sum_logprobs / cur_len ** self.length_penalty


def calc_banned_ngram_tokens(prev_input_ids, num_hypos, no_repeat_ngram_size, cur_len):
    """generation_utils.py:848-868 (token lists already on the host)."""
    if cur_len + 1 < no_repeat_ngram_size:
        return [[] for _ in range(num_hypos)]
    banned = []
    for idx in range(num_hypos):
        gen = prev_input_ids[idx]
        seen = {}
        for i in range(len(gen) - no_repeat_ngram_size + 1):
            ng = tuple(gen[i:i + no_repeat_ngram_size])
            seen.setdefault(ng[:-1], []).append(ng[-1])
        start = cur_len + 1 - no_repeat_ngram_size
        banned.append(seen.get(tuple(gen[start:cur_len]), []))
    return banned


class Generator:
    def __init__(self, model):
        self.model = model
        self.eng = None

    # ------------------------------------------------------------------ memory
    @torch.no_grad()
    def encode(self, reviews, reviews_mask, field, field_value, img, img_mask, num_beams):
        eng = self.eng = self.model._ensure_engine(reviews.device)
        cfg = eng.cfg
        if cfg.dataset == "text":
            raise NotImplementedError("generation is implemented for the multimodal model")
        dev = reviews.device
        D, H = cfg.d_model, cfg.heads
        B, R, S = reviews.shape
        if S > 208:
            raise ValueError("review frames up to 208 tokens are supported")
        Sp = 128 * math.ceil(S / 128)
        F = 47 if cfg.dataset == "yelp" else 133
        n_img, ik = img.shape[1], img.shape[2]
        T = B * R * Sp
        Tm = T + B * F + B * n_img * ik
        eng.refresh_bf16_weights()
        bf = lambda *s: torch.empty(s, device=dev, dtype=torch.bfloat16)
        f32 = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        g = ops.gemm
        bm = "bart_model.model."
        # frames padded to a multiple of 128 (pad id 1, invalid)
        ids = torch.ones(B * R, Sp, device=dev, dtype=torch.int32)
        ids[:, :S] = reviews.reshape(B * R, S).to(torch.int32)
        valid = torch.zeros(B * R, Sp, device=dev, dtype=torch.uint8)
        valid[:, :S] = (reviews_mask.reshape(B * R, S) != 0).to(torch.uint8)
        MEM = bf(Tm, D)
        # table + image memories
        t = "table_encoder."
        W1n = t + ("hours_embedding.weight" if cfg.dataset == "yelp" else "price_embedding.weight")
        W0, W1 = (eng.w32(t + "rating_embedding.weight"), eng.w32(W1n)) if cfg.dataset == "yelp" else (eng.w32(W1n), eng.w32(t + "rating_embedding.weight"))
        tabX, tab_valid, tab_h = bf(B * F, 2 * D), torch.zeros(B, F, device=dev, dtype=torch.uint8), bf(B * F, D)
        ops.table_fwd(cfg.dataset, B, eng.w32(bm + "shared.weight"), field, list(field_value), W0, W1, tabX, tab_valid)
        g(tabX, eng.w16(t + "fc.weight"), tab_h, bias=eng.w32(t + "fc.bias"), act=ops.ACT_RELU)
        g(tab_h, eng.w16(t + "linear.weight"), MEM[T:T + B * F])
        feats = img.reshape(B * n_img * ik, 1024)
        if feats.dtype == torch.float32:
            feats = ops.cast_bf16(feats.contiguous(), bf(B * n_img * ik, 1024))
        g(feats, eng.w16("img_encoder.linear.weight"), MEM[T + B * F:])
        # encoder (eval: no dropout); a frame of Sp rows is Sp/128 query tiles over one key entity of <= 208 rows
        tiles = Sp // 128
        N = B * R
        pre = bm + "encoder."
        x, x1, nxt = bf(T, D), bf(T, D), bf(T, D)
        qkv, ctx, o, a_buf, f_buf = bf(T, 3 * D), bf(T, D), bf(T, D), bf(T, cfg.ffn_dim), bf(T, D)
        mean, rstd = f32(T), f32(T)
        lse = f32(N * tiles, H, 1, 128)
        kvalid = valid.reshape(-1)
        ops.embed_ln_fwd(ids.reshape(-1), eng.w32(bm + "shared.weight"), eng.w32(pre + "embed_positions.weight"), None, None,
                         eng.w32(pre + "layernorm_embedding.weight"), eng.w32(pre + "layernorm_embedding.bias"), x, mean, rstd,
                         T, Sp, 0.0, 0, 0)
        Sk = min(Sp, 208)
        for l in range(cfg.encoder_layers):
            lp = pre + "layers.%d." % l
            s_ = lp + "self_attn."
            g(x, eng.w16(s_ + "q_proj.weight", s_ + "v_proj.weight"), qkv, bias=eng.w32(s_ + "q_proj.bias", s_ + "v_proj.bias"))
            ops.attn_fwd(ops.attn_args(Q=qkv, ldq=3 * D, q_col=0, KV=qkv, ldkv=3 * D, k_col=D, v_col=2 * D, O=ctx, ldo=D, LSE=lse,
                                       key_valid=kvalid, ent_valid=None, inv_n=None, n_qseq=N * tiles, H=H, R=tiles, causal=0,
                                       E_total=1, scale=cfg.head_dim ** -0.5, mods=[(0, 0, 1, Sk, 0, 0, Sp)]))
            g(ctx, eng.w16(s_ + "out_proj.weight"), o, bias=eng.w32(s_ + "out_proj.bias"))
            ops.add_ln_fwd(x, o, eng.w32(lp + "self_attn_layer_norm.weight"), eng.w32(lp + "self_attn_layer_norm.bias"), x1, mean, rstd, 0.0, 0, 0)
            g(x1, eng.w16(lp + "fc1.weight"), a_buf, bias=eng.w32(lp + "fc1.bias"), act=ops.ACT_GELU)
            g(a_buf, eng.w16(lp + "fc2.weight"), f_buf, bias=eng.w32(lp + "fc2.bias"))
            out = MEM[:T] if l == cfg.encoder_layers - 1 else nxt
            ops.add_ln_fwd(x1, f_buf, eng.w32(lp + "final_layer_norm.weight"), eng.w32(lp + "final_layer_norm.bias"), out, mean, rstd, 0.0, 0, 0)
            x, nxt = out, x
        # validity bookkeeping (no leave-one-out at test time: every review is a source)
        img_u8 = img_mask.to(torch.uint8)
        mem_valid = torch.cat([kvalid, tab_valid.reshape(-1), img_u8.repeat_interleave(ik, dim=1).reshape(-1)]).contiguous()
        text_ent = valid.reshape(B, R, Sp).amax(dim=2)
        tab_ent = tab_valid.amax(dim=1, keepdim=True)
        ent_valid = torch.cat([text_ent, tab_ent, img_u8], dim=1).contiguous()
        cnt = lambda m: m.sum(dim=1).float()
        inv = lambda c: torch.where(c > 0, 1.0 / c.clamp(min=1), torch.zeros_like(c))
        inv_n_biz = torch.stack([inv(cnt(text_ent)), inv(cnt(tab_ent)), inv(cnt(img_u8))], dim=1)          # [B,3]
        pres = torch.stack([tab_ent[:, 0], (img_u8.sum(dim=1) > 0).to(torch.uint8)], dim=1).contiguous()   # [B,2]
        # cross-attention K|V of the memory: once per decoder layer, un-expanded per business
        kv = []
        for l in range(cfg.decoder_layers):
            c = bm + "decoder.layers.%d.encoder_attn." % l
            kv.append(g(MEM, eng.w16(c + "k_proj.weight", c + "v_proj.weight"), bf(Tm, 2 * D),
                        bias=eng.w32(c + "k_proj.bias", c + "v_proj.bias")))
        return dict(B=B, R=R, Sp=Sp, Sk=Sk, F=F, n_img=n_img, ik=ik, T=T, Tm=Tm, kv=kv, mem_valid=mem_valid, ent_valid=ent_valid,
                    inv_n=inv_n_biz.repeat_interleave(num_beams, dim=0).contiguous(), inv_n_biz=inv_n_biz.contiguous(), pres=pres,
                    beams=num_beams, ws=None, cws=None)

    # ------------------------------------------------------------------ one decode step (all beams)
    @torch.no_grad()
    def last_logits(self, st, input_ids, rating_diff):
        """input_ids [N, cur_len] (N = B*beams, beams of a business adjacent) -> fp32 logits [N, V] of the last position."""
        eng, cfg = self.eng, self.eng.cfg
        dev = input_ids.device
        D, H, V = cfg.d_model, cfg.heads, cfg.vocab_size
        N, cur = input_ids.shape
        S = 128
        if cur > S:
            raise ValueError("decoder frames up to 128 tokens")
        T = N * S
        beams, B = st["beams"], st["B"]
        Et = st["R"] + 1 + st["n_img"]
        if st["ws"] is None:
            bf = lambda *s: torch.empty(s, device=dev, dtype=torch.bfloat16)
            f32 = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
            st["ws"] = dict(x=bf(T, D), x1=bf(T, D), x2=bf(T, D), nxt=bf(T, D), qkv=bf(T, 3 * D), ctx=bf(T, D), o=bf(T, D), qc=bf(T, D),
                            A3=bf(3, T, D), O3=bf(3, T, D), U=bf(2, T, D), AB=bf(2, T, D), yc=bf(T, D), a=bf(T, cfg.ffn_dim), f=bf(T, D),
                            mean=f32(T), rstd=f32(T), lse=f32(N, H, 1, S), lse_c=f32(N, H, Et, S),
                            ids=torch.ones(N, S, device=dev, dtype=torch.int32), logits=f32(N, (V + 3) // 4 * 4)[:, :V])   # 16-byte row pitch for TMA
        w = st["ws"]
        g = ops.gemm
        bm = "bart_model.model."
        pre = bm + "decoder."
        w["ids"].fill_(cfg.pad_token_id)
        w["ids"][:, :cur] = input_ids.to(torch.int32)
        x = w["x"]
        ops.embed_ln_fwd(w["ids"].reshape(-1), eng.w32(bm + "shared.weight"), eng.w32(pre + "embed_positions.weight"),
                         rating_diff.reshape(-1).float().contiguous(), eng.w32(pre + "rating_embeddings"),
                         eng.w32(pre + "layernorm_embedding.weight"), eng.w32(pre + "layernorm_embedding.bias"), x, w["mean"], w["rstd"],
                         T, S, 0.0, 0, 0)
        R, Sp, Sk, F, n_img, ik, Tt = st["R"], st["Sp"], st["Sk"], st["F"], st["n_img"], st["ik"], st["T"]
        mods = [(0, 0, R, Sk, 0, 0, Sp), (Tt, T * D, 1, F, 0, R, 0), (Tt + B * F, 2 * T * D, n_img, ik, 0, R + 1, 0)]
        nxt = w["nxt"]
        for l in range(cfg.decoder_layers):
            lp = pre + "layers.%d." % l
            s_, c = lp + "self_attn.", lp + "encoder_attn."
            g(x, eng.w16(s_ + "q_proj.weight", s_ + "v_proj.weight"), w["qkv"], bias=eng.w32(s_ + "q_proj.bias", s_ + "v_proj.bias"))
            ops.attn_fwd(ops.attn_args(Q=w["qkv"], ldq=3 * D, q_col=0, KV=w["qkv"], ldkv=3 * D, k_col=D, v_col=2 * D, O=w["ctx"], ldo=D,
                                       LSE=w["lse"], key_valid=None, ent_valid=None, inv_n=None, n_qseq=N, H=H, R=1, causal=1, E_total=1,
                                       scale=cfg.head_dim ** -0.5, mods=[(0, 0, 1, S, 0, 0)]))
            g(w["ctx"], eng.w16(s_ + "out_proj.weight"), w["o"], bias=eng.w32(s_ + "out_proj.bias"))
            ops.add_ln_fwd(x, w["o"], eng.w32(lp + "self_attn_layer_norm.weight"), eng.w32(lp + "self_attn_layer_norm.bias"), w["x1"],
                           w["mean"], w["rstd"], 0.0, 0, 0)
            g(w["x1"], eng.w16(c + "q_proj.weight"), w["qc"], bias=eng.w32(c + "q_proj.bias"))
            ops.attn_fwd(ops.attn_args(Q=w["qc"], ldq=D, q_col=0, KV=st["kv"][l], ldkv=2 * D, k_col=0, v_col=D, O=w["A3"], ldo=D, LSE=w["lse_c"],
                                       key_valid=st["mem_valid"], ent_valid=st["ent_valid"], inv_n=st["inv_n"], n_qseq=N, H=H, R=beams, causal=0,
                                       E_total=Et, scale=cfg.head_dim ** -0.5, mods=mods))
            g(w["A3"].view(3 * T, D), eng.w16(c + "out_proj.weight"), w["O3"].view(3 * T, D), bias=eng.w32(c + "out_proj.bias"))
            ops.gemm_cat(w["O3"][0], w["O3"][1], eng.w16(c + "alpha_proj.weight"), w["U"][0], bias=eng.w32(c + "alpha_proj.bias"))
            ops.gemm_cat(w["O3"][0], w["O3"][2], eng.w16(c + "beta_proj.weight"), w["U"][1], bias=eng.w32(c + "beta_proj.bias"))
            ops.gate_fwd(w["O3"], w["U"], st["pres"], w["yc"], w["AB"], T, beams * S, D)
            ops.add_ln_fwd(w["x1"], w["yc"], eng.w32(lp + "encoder_attn_layer_norm.weight"), eng.w32(lp + "encoder_attn_layer_norm.bias"),
                           w["x2"], w["mean"], w["rstd"], 0.0, 0, 0)
            g(w["x2"], eng.w16(lp + "fc1.weight"), w["a"], bias=eng.w32(lp + "fc1.bias"), act=ops.ACT_GELU)
            g(w["a"], eng.w16(lp + "fc2.weight"), w["f"], bias=eng.w32(lp + "fc2.bias"))
            ops.add_ln_fwd(w["x2"], w["f"], eng.w32(lp + "final_layer_norm.weight"), eng.w32(lp + "final_layer_norm.bias"), nxt,
                           w["mean"], w["rstd"], 0.0, 0, 0)
            x, nxt = nxt, x
        last = x.view(N, S, D)[:, cur - 1]        # strided [N, D] view: the GEMM reads it in place
        g(last, eng.w16(bm + "shared.weight"), w["logits"], bias=eng.w32_flb())
        return w["logits"]

    # ------------------------------------------------------------------ one cached decode step (all beams)
    @torch.no_grad()
    def step_logits(self, st, input_ids, rating_diff):
        """Incremental decoding (`_use_saved_state` / cached branch of `get_head_output`, modeling_multimodalsum.py:774-815,
        889-920): only the newest token of every hypothesis goes through the decoder.  Self-attention K|V of all earlier
        positions live in per-layer caches [N, 128, 2D] (the new row is written in place by the K|V GEMM), cross-attention
        K|V are the static per-business projections from `encode`.  Call `reorder_cache(st, beam_idx)` after every beam
        re-ranking (`_reorder_cache`, :3103-3115).  input_ids [N, cur_len] -> fp32 logits [N, V] of the last position.

        The attention kernels work on 128-row query tiles: the self-attention query of hypothesis n sits in row t of its own
        causal frame (rows != t are ignored), and the `beams` cross-attention queries of a business share one frame (rows
        0..beams-1) because they attend to the same memory."""
        eng, cfg = self.eng, self.eng.cfg
        dev = input_ids.device
        D, H, V, FF = cfg.d_model, cfg.heads, cfg.vocab_size, cfg.ffn_dim
        N, cur = input_ids.shape
        S = 128
        if cur > S:
            raise ValueError("decoder frames up to 128 tokens")
        if st["beams"] > S:
            raise ValueError("at most 128 beams")
        t = cur - 1
        beams, B = st["beams"], st["B"]
        Et = st["R"] + 1 + st["n_img"]
        if st["cws"] is None:
            bf = lambda *s: torch.empty(s, device=dev, dtype=torch.bfloat16)
            zbf = lambda *s: torch.zeros(s, device=dev, dtype=torch.bfloat16)
            f32 = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
            L = cfg.decoder_layers
            st["cws"] = dict(x=bf(N, D), x1=bf(N, D), x2=bf(N, D), nxt=bf(N, D), o=bf(N, D), qc=bf(N, D), a=bf(N, FF), f=bf(N, D),
                             qf=zbf(N * S, D), ctx=zbf(N * S, D),                 # self-attention query / context frames
                             kvs=[zbf(N, S, 2 * D) for _ in range(L)],            # self-attention K|V caches (finite everywhere)
                             kvs_alt=[zbf(N, S, 2 * D) for _ in range(L)],
                             qcf=zbf(B * S, D), A3f=zbf(3, B * S, D),             # cross-attention frames: one per business
                             A3=bf(3, N, D), O3=bf(3, N, D), U=bf(2, N, D), AB=bf(2, N, D), yc=bf(N, D),
                             mean=f32(N), rstd=f32(N), lse=f32(N, H, 1, S), lse_c=f32(B, H, Et, S),
                             ids=torch.empty(N, device=dev, dtype=torch.int32),
                             logits=f32(N, (V + 3) // 4 * 4)[:, :V], pos=-1)
        w = st["cws"]
        if t != w["pos"] + 1:
            raise ValueError("step_logits must be called with consecutive lengths (got position %d after %d)" % (t, w["pos"]))
        w["pos"] = t
        g = ops.gemm
        bm = "bart_model.model."
        pre = bm + "decoder."
        w["ids"].copy_(input_ids[:, t])
        x = w["x"]
        # position t for every row: the kernel adds P[(row % S) + 2], so pass S = 1 and the table shifted by t rows
        ops.embed_ln_fwd(w["ids"], eng.w32(bm + "shared.weight"), eng.w32(pre + "embed_positions.weight")[t:],
                         rating_diff.reshape(-1).float().contiguous(), eng.w32(pre + "rating_embeddings"),
                         eng.w32(pre + "layernorm_embedding.weight"), eng.w32(pre + "layernorm_embedding.bias"), x, w["mean"], w["rstd"],
                         N, 1, 0.0, 0, 0)
        R, Sp, Sk, F, n_img, ik, Tt = st["R"], st["Sp"], st["Sk"], st["F"], st["n_img"], st["ik"], st["T"]
        Tf = B * S
        mods = [(0, 0, R, Sk, 0, 0, Sp), (Tt, Tf * D, 1, F, 0, R, 0), (Tt + B * F, 2 * Tf * D, n_img, ik, 0, R + 1, 0)]
        q_row = w["qf"].view(N, S, D)[:, t]          # strided [N, D] views: the GEMMs read / write them in place
        ctx_row = w["ctx"].view(N, S, D)[:, t]
        nxt = w["nxt"]
        for l in range(cfg.decoder_layers):
            lp = pre + "layers.%d." % l
            s_, c = lp + "self_attn.", lp + "encoder_attn."
            wqkv, bqkv = eng.w16(s_ + "q_proj.weight", s_ + "v_proj.weight"), eng.w32(s_ + "q_proj.bias", s_ + "v_proj.bias")
            cache = w["kvs"][l]
            g(x, wqkv[:D], q_row, bias=bqkv[:D])
            g(x, wqkv[D:], cache[:, t], bias=bqkv[D:])                      # appends this position's K|V to the cache
            ops.attn_fwd(ops.attn_args(Q=w["qf"], ldq=D, q_col=0, KV=cache.view(N * S, 2 * D), ldkv=2 * D, k_col=0, v_col=D, O=w["ctx"],
                                       ldo=D, LSE=w["lse"], key_valid=None, ent_valid=None, inv_n=None, n_qseq=N, H=H, R=1, causal=1,
                                       E_total=1, scale=cfg.head_dim ** -0.5, mods=[(0, 0, 1, S, 0, 0)]))
            g(ctx_row, eng.w16(s_ + "out_proj.weight"), w["o"], bias=eng.w32(s_ + "out_proj.bias"))
            ops.add_ln_fwd(x, w["o"], eng.w32(lp + "self_attn_layer_norm.weight"), eng.w32(lp + "self_attn_layer_norm.bias"), w["x1"],
                           w["mean"], w["rstd"], 0.0, 0, 0)
            g(w["x1"], eng.w16(c + "q_proj.weight"), w["qc"], bias=eng.w32(c + "q_proj.bias"))
            w["qcf"].view(B, S, D)[:, :beams].copy_(w["qc"].view(B, beams, D))
            ops.attn_fwd(ops.attn_args(Q=w["qcf"], ldq=D, q_col=0, KV=st["kv"][l], ldkv=2 * D, k_col=0, v_col=D, O=w["A3f"], ldo=D,
                                       LSE=w["lse_c"], key_valid=st["mem_valid"], ent_valid=st["ent_valid"], inv_n=st["inv_n_biz"], n_qseq=B,
                                       H=H, R=1, causal=0, E_total=Et, scale=cfg.head_dim ** -0.5, mods=mods))
            w["A3"].view(3, B, beams, D).copy_(w["A3f"].view(3, B, S, D)[:, :, :beams])
            g(w["A3"].view(3 * N, D), eng.w16(c + "out_proj.weight"), w["O3"].view(3 * N, D), bias=eng.w32(c + "out_proj.bias"))
            ops.gemm_cat(w["O3"][0], w["O3"][1], eng.w16(c + "alpha_proj.weight"), w["U"][0], bias=eng.w32(c + "alpha_proj.bias"))
            ops.gemm_cat(w["O3"][0], w["O3"][2], eng.w16(c + "beta_proj.weight"), w["U"][1], bias=eng.w32(c + "beta_proj.bias"))
            ops.gate_fwd(w["O3"], w["U"], st["pres"], w["yc"], w["AB"], N, beams, D)
            ops.add_ln_fwd(w["x1"], w["yc"], eng.w32(lp + "encoder_attn_layer_norm.weight"), eng.w32(lp + "encoder_attn_layer_norm.bias"),
                           w["x2"], w["mean"], w["rstd"], 0.0, 0, 0)
            g(w["x2"], eng.w16(lp + "fc1.weight"), w["a"], bias=eng.w32(lp + "fc1.bias"), act=ops.ACT_GELU)
            g(w["a"], eng.w16(lp + "fc2.weight"), w["f"], bias=eng.w32(lp + "fc2.bias"))
            ops.add_ln_fwd(w["x2"], w["f"], eng.w32(lp + "final_layer_norm.weight"), eng.w32(lp + "final_layer_norm.bias"), nxt,
                           w["mean"], w["rstd"], 0.0, 0, 0)
            x, nxt = nxt, x
        w["x"], w["nxt"] = x, nxt
        g(x, eng.w16(bm + "shared.weight"), w["logits"], bias=eng.w32_flb())
        return w["logits"]

    @torch.no_grad()
    def reorder_cache(self, st, beam_idx):
        """`_reorder_cache` (:3103-3115): hypothesis i continues hypothesis beam_idx[i]; only the self-attention caches move
        (the cross-attention K|V are per business and beam_idx never crosses businesses)."""
        w = st["cws"]
        if w is None:
            return
        for l in range(len(w["kvs"])):
            torch.index_select(w["kvs"][l], 0, beam_idx, out=w["kvs_alt"][l])
            w["kvs"][l], w["kvs_alt"][l] = w["kvs_alt"][l], w["kvs"][l]

    # ------------------------------------------------------------------ public entry point (src/test.py:152-158)
    @torch.no_grad()
    def generate(self, reviews, reviews_mask, field, field_value, img, img_mask, rating_diff=None, num_beams=4, max_length=20,
                 min_length=0, length_penalty=1.0, no_repeat_ngram_size=3, early_stopping=True, use_cache=True):
        """use_cache=True: incremental decoding with self-attention K|V caches (`step_logits`); False: recompute the whole
        prefix every step (`last_logits`, kept as the cross-check of the cached path)."""
        cfg = self.model.cfg
        B = reviews.shape[0]
        st = self.encode(reviews, reviews_mask, field, field_value, img, img_mask, num_beams)
        rd = torch.zeros(B, device=reviews.device) if rating_diff is None else rating_diff.reshape(B).float()
        rd = rd.repeat_interleave(num_beams).contiguous()
        if use_cache:
            logits_fn, reorder_fn = (lambda ids: self.step_logits(st, ids, rd)), (lambda beam_idx: self.reorder_cache(st, beam_idx))
        else:
            logits_fn, reorder_fn = (lambda ids: self.last_logits(st, ids, rd)), None
        return beam_search(logits_fn, B, cfg.vocab_size, reviews.device, num_beams=num_beams,
                           max_length=max_length, min_length=min_length, length_penalty=length_penalty,
                           no_repeat_ngram_size=no_repeat_ngram_size, early_stopping=early_stopping, pad=cfg.pad_token_id,
                           bos=cfg.bos_token_id, eos=cfg.eos_token_id, reorder_fn=reorder_fn)


def beam_search(logits_fn, B, V, dev, num_beams=4, max_length=20, min_length=0, length_penalty=1.0, no_repeat_ngram_size=3,
                early_stopping=True, pad=1, bos=0, eos=2, reorder_fn=None):
    """_generate_beam_search (modeling_multimodalsum.py:2803-3067), do_sample=False.  `logits_fn(input_ids[N, cur_len])`
    returns the fp32 next-token logits [N, V] of the last position (N = B*num_beams, beams of a business adjacent)."""
    N = B * num_beams
    input_ids = torch.full((N, 1), eos, dtype=torch.long, device=dev)     # decoder_start_token_id = 2 (cfg/bart-large.json)
    hyps = [BeamHypotheses(num_beams, max_length, length_penalty, early_stopping) for _ in range(B)]
    beam_scores = torch.zeros(B, num_beams, device=dev)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    done = [False] * B
    cur_len = 1
    next_scores = next_tokens = None
    while cur_len < max_length:
        logits = logits_fn(input_ids).clone()
        # adjust_logits_during_generation (:3084-3089)
        if cur_len == 1:
            keep = logits[:, bos].clone(); logits.fill_(float("-inf")); logits[:, bos] = keep
        if cur_len == max_length - 1:
            keep = logits[:, eos].clone(); logits.fill_(float("-inf")); logits[:, eos] = keep
        scores = torch.log_softmax(logits, dim=-1)
        # postprocess_next_token_scores (generation_utils.py:57-99)
        if cur_len < min_length:
            scores[:, eos] = float("-inf")
        ids_host = input_ids.tolist()
        if no_repeat_ngram_size > 0:
            for i, banned in enumerate(calc_banned_ngram_tokens(ids_host, N, no_repeat_ngram_size, cur_len)):
                if banned:
                    scores[i, banned] = float("-inf")
        nxt = (scores + beam_scores[:, None]).view(B, num_beams * V)
        next_scores, next_tokens = torch.topk(nxt, 2 * num_beams, dim=1, largest=True, sorted=True)
        ns_host, nt_host = next_scores.tolist(), next_tokens.tolist()
        next_batch_beam = []
        for b in range(B):
            if done[b]:
                next_batch_beam.extend([(0.0, pad, 0)] * num_beams)
                continue
            sent = []
            for rank, (tok_id, tok_score) in enumerate(zip(nt_host[b], ns_host[b])):
                beam_id, token_id = tok_id // V, tok_id % V
                eff = b * num_beams + beam_id
                if token_id == eos:
                    if rank >= num_beams:
                        continue
                    hyps[b].add(list(ids_host[eff]), tok_score)
                else:
                    sent.append((tok_score, token_id, eff))
                if len(sent) == num_beams:
                    break
            done[b] = done[b] or hyps[b].is_done(max(ns_host[b]), cur_len)
            assert len(sent) == num_beams, "Beam should always be full"
            next_batch_beam.extend(sent)
        if all(done):
            break
        beam_scores = torch.tensor([x[0] for x in next_batch_beam], device=dev, dtype=torch.float32)
        beam_tokens = torch.tensor([x[1] for x in next_batch_beam], device=dev, dtype=torch.long)
        beam_idx = torch.tensor([x[2] for x in next_batch_beam], device=dev, dtype=torch.long)
        input_ids = torch.cat([input_ids[beam_idx, :], beam_tokens.unsqueeze(1)], dim=-1)
        cur_len += 1
        # the reference re-gathers memories and caches with beam_idx here (:2957, _reorder_cache); the per-business memory is
        # un-expanded (beam_idx never crosses businesses), so only an incremental decoder's self-attention caches move
        if reorder_fn is not None:
            reorder_fn(beam_idx)
    ids_host = input_ids.tolist()
    bs_host = beam_scores.tolist()
    for b in range(B):
        if done[b]:
            continue
        for k in range(num_beams):
            eff = b * num_beams + k
            hyps[b].add(list(ids_host[eff]), bs_host[eff])
    best = [sorted(h.beams, key=lambda x: x[0])[-1][1] for h in hyps]
    lens = [len(h) for h in best]
    if min(lens) != max(lens):
        width = min(max(lens) + 1, max_length)
        out = torch.full((B, width), pad, dtype=torch.long)
        for i, h in enumerate(best):
            out[i, :lens[i]] = torch.tensor(h)
            if lens[i] < max_length:
                out[i, lens[i]] = eos
    else:
        out = torch.tensor(best, dtype=torch.long)
    return out.to(dev)
