"""Repeatability of generate() (cached graph path vs use_cache=False) at the config-5 test shape, with the allocator cache and
every SM's on-chip memory poisoned with NaN patterns before each phase: prints the token agreement and the output shapes."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from multimodalsum_b200 import ops
from multimodalsum_b200.generation import Generator
from multimodalsum_b200.modules import MultimodalSum, YelpTableEncoder
from multimodalsum_b200.synth import ModelConfig, make_batch, make_state_dict

poison = len(sys.argv) > 1 and sys.argv[1] == "poison"
B, beams = 64, 4
cfg = ModelConfig(dataset="yelp", encoder_layers=2, decoder_layers=2, dropout=0.0)
sd = make_state_dict(cfg, seed=41, gates_open=True, logits_bias_std=1.0)
batch = make_batch(cfg, B, seed=42, n_reviews=8, seq_len=158, len_range=(100, 150)).to("cuda")


def poison_all():
    if not poison:
        return
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    junk = [torch.full((n,), float("nan"), device="cuda") for n in (1 << 29, 1 << 28, 1 << 26, 1 << 26, 1 << 24, 1 << 24, 1 << 22, 1 << 20)]
    junk += [torch.full((1 << 16,), float("nan"), device="cuda") for _ in range(64)]
    del junk
    ops.debug_poison()
    torch.cuda.synchronize()


outs = []
for rep in range(3):
    poison_all()
    model = MultimodalSum(TableEncoder=YelpTableEncoder, config=cfg)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    gen = Generator(model)
    args = (batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask)
    poison_all()
    out = gen.generate(*args, num_beams=beams, max_length=12, no_repeat_ngram_size=3, early_stopping=True)
    poison_all()
    out2 = gen.generate(*args, num_beams=beams, max_length=12, no_repeat_ngram_size=3, early_stopping=True)
    poison_all()
    ref = gen.generate(*args, num_beams=beams, max_length=12, no_repeat_ngram_size=3, early_stopping=True, use_cache=False)
    W = max(out.shape[1], ref.shape[1])
    padw = lambda t: torch.nn.functional.pad(t, (0, W - t.shape[1]), value=cfg.pad_token_id)
    print("rep", rep, "shapes", tuple(out.shape), tuple(out2.shape), tuple(ref.shape), "cached==cached2", torch.equal(out, out2),
          "agreement cached vs recompute %.4f" % float((padw(out) == padw(ref)).float().mean()), flush=True)
    outs.append((out.clone(), ref.clone()))
    del gen, model
print("cached identical across reps:", all(o[0].shape == outs[0][0].shape and torch.equal(o[0], outs[0][0]) for o in outs),
      " recompute identical across reps:", all(o[1].shape == outs[0][1].shape and torch.equal(o[1], outs[0][1]) for o in outs))
