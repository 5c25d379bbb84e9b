"""BASELINE config 5 shape: beam-4 generation, 64 businesses, 8 reviews x 158 tokens + 47 table fields + 10x196 image keys,
BART-large random init.  Reports encode time and time per decode step for the prefix-recompute path and for incremental decoding."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalsum_b200.generation import Generator
from multimodalsum_b200.modules import MultimodalSum, YelpTableEncoder
from multimodalsum_b200.synth import ModelConfig, make_batch
B = int(os.environ.get("MMSUM_B", "64")); beams = 4; steps = int(os.environ.get("MMSUM_STEPS", "8"))
cfg = ModelConfig(dataset="yelp", dropout=0.0)
torch.manual_seed(0)
model = MultimodalSum(TableEncoder=YelpTableEncoder, config=cfg).cuda().eval()
batch = make_batch(cfg, B, seed=5, n_reviews=8, seq_len=158, fixed_len=150, n_valid_imgs=10).to("cuda")
gen = Generator(model)
torch.cuda.synchronize(); t0 = time.time()
st = gen.encode(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask, beams)
torch.cuda.synchronize(); t_enc = time.time() - t0
N = B * beams
rd = torch.zeros(N, device="cuda")
ids = torch.randint(3, cfg.vocab_size, (N, 1 + steps + 2), device="cuda"); ids[:, 0] = 2
for cur in (1, 2):
    gen.last_logits(st, ids[:, :cur].contiguous(), rd)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for cur in range(3, 3 + steps):
    gen.last_logits(st, ids[:, :cur].contiguous(), rd)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / steps
print("config-5 shape: B=%d beams=%d  encode (memory + 12 cached cross K|V) %.1f ms;  decode step %.1f ms (%d hypotheses) -> %.0f businesses*tokens/s; "
      "max mem %.1f GB" % (B, beams, t_enc * 1e3, ms, N, B / (ms / 1e3), torch.cuda.max_memory_allocated() / 2**30))
# incremental decoding: one token per hypothesis, self-attention K|V caches, beam re-ordering every step
st2 = gen.encode(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask, beams)
perm = torch.arange(N, device="cuda")
for cur in (1, 2):
    gen.step_logits(st2, ids[:, :cur].contiguous(), rd); gen.reorder_cache(st2, perm)
torch.cuda.synchronize()
s.record()
for cur in range(3, 3 + steps):
    gen.step_logits(st2, ids[:, :cur].contiguous(), rd); gen.reorder_cache(st2, perm)
e.record(); torch.cuda.synchronize()
ms2 = s.elapsed_time(e) / steps
print("cached decode step (incl. cache re-ordering) %.1f ms -> %.0f businesses*tokens/s (%.1fx)" % (ms2, B / (ms2 / 1e3), ms / ms2))
t0 = time.time()
out = gen.generate(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask, num_beams=beams, max_length=12)
torch.cuda.synchronize()
print("generate(max_length=12): %.2f s, output %s" % (time.time() - t0, tuple(out.shape)))
