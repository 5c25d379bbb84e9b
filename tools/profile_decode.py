"""One profiled decode step at the BASELINE config-5 shape (64 businesses x 4 beams, 8 x 158-token reviews, 47 fields, 10 x 196
image keys, BART-large), launched kernel by kernel (no CUDA graph) inside a cudaProfilerStart/Stop window:
   MMSUM_DECODE_GRAPH=0 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
       --log-file gpurun_out/decode_launches.csv python tools/profile_decode.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalsum_b200.generation import Generator
from multimodalsum_b200.modules import MultimodalSum, YelpTableEncoder
from multimodalsum_b200.synth import ModelConfig, make_batch

B, beams = int(os.environ.get("MMSUM_B", "64")), 4
cfg = ModelConfig(dataset="yelp", dropout=0.0)
torch.manual_seed(0)
model = MultimodalSum(TableEncoder=YelpTableEncoder, config=cfg).cuda().eval()
b = make_batch(cfg, B, seed=5, n_reviews=8, seq_len=158, fixed_len=150, n_valid_imgs=10).to("cuda")
gen = Generator(model)
st = gen.encode(b.reviews, b.reviews_mask, b.field, b.field_value, b.img, b.img_mask, beams)
N = B * beams
rd = torch.zeros(N, device="cuda")
ids = torch.randint(3, cfg.vocab_size, (N, 40), device="cuda")
perm = torch.arange(N, device="cuda")
for cur in range(1, 21):
    gen.step_logits(st, ids[:, :cur], rd)
    gen.reorder_cache(st, perm)
torch.cuda.synchronize()
torch.cuda.profiler.start()
gen.step_logits(st, ids[:, :21], rd)
gen.reorder_cache(st, perm)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for cur in range(22, 30):
    gen.step_logits(st, ids[:, :cur], rd)
    gen.reorder_cache(st, perm)
e.record()
torch.cuda.synchronize()
print("decode step %.3f ms (%s)" % (s.elapsed_time(e) / 8, "graph" if st.cws["graph"] else "kernel by kernel"))
