"""One profiled training step (config 2 shape) for ncu: warm-up steps run outside the cudaProfilerStart/Stop window.
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/profile_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalsum_b200.modules import MultimodalSum, YelpTableEncoder
from multimodalsum_b200.synth import ModelConfig, make_batch

B = int(os.environ.get("MMSUM_B", "16"))
steps = int(os.environ.get("MMSUM_PROFILE_STEPS", "1"))
cfg = ModelConfig(dataset="yelp", dropout=0.1)
torch.manual_seed(0)
model = MultimodalSum(TableEncoder=YelpTableEncoder, config=cfg).cuda().train()
b = make_batch(cfg, B, seed=1234, fixed_len=100, n_valid_imgs=10).to("cuda")

from multimodalsum_b200.optim import get_optimizer
opt = get_optimizer(model.engine if hasattr(model, "engine") and model.engine is not None else model._ensure_engine(torch.device("cuda")),
                    3e-5, ["bias", "LayerNorm.weight"], list(model.named_parameters()), None, max_grad_norm=1.0)

def step():
    loss = model(b.reviews, b.reviews_mask, b.reviews_rating, b.field, b.field_value, b.img, b.img_mask)[0]
    model.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss

for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(steps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
