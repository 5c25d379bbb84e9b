"""The config-5 generation test's sequence under a NaN-poisoned allocator cache: is generate() repeatable (call 1 = first token
eager + recorded graph, call 2 = graph from the first token)?  argv: poison kinds 'mem', 'chip', 'both', 'none'."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from multimodalsum_b200 import ops
from multimodalsum_b200.generation import Generator
from multimodalsum_b200.modules import MultimodalSum, YelpTableEncoder
from multimodalsum_b200.synth import ModelConfig, make_batch, make_state_dict
kind = sys.argv[1] if len(sys.argv) > 1 else "both"


def poison_all():
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    if kind in ("mem", "both"):
        junk = [torch.full((n,), float("nan"), device="cuda") for n in (1 << 29, 1 << 28, 1 << 27, 1 << 26, 1 << 26, 1 << 24, 1 << 24, 1 << 22, 1 << 20)]
        junk += [torch.full((1 << 16,), float("nan"), device="cuda") for _ in range(64)]
        junk += [torch.full((1 << 12,), float("nan"), device="cuda") for _ in range(256)]
        del junk
    if kind in ("chip", "both"):
        ops.debug_poison()
    torch.cuda.synchronize()


B, beams = 64, 4
cfg = ModelConfig(dataset="yelp", encoder_layers=2, decoder_layers=2, dropout=0.0)
sd = make_state_dict(cfg, seed=41, gates_open=True, logits_bias_std=1.0)
batch = make_batch(cfg, B, seed=42, n_reviews=8, seq_len=158, len_range=(100, 150)).to("cuda")
poison_all()
model = MultimodalSum(TableEncoder=YelpTableEncoder, config=cfg)
model.load_state_dict(sd, strict=False)
model = model.cuda().eval()
gen = Generator(model)
args = (batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask)
st = gen.encode(*args, beams)
N = B * beams
rd = torch.zeros(N, device="cuda")
g = torch.Generator().manual_seed(9)
ids = torch.cat([torch.full((N, 1), cfg.eos_token_id), torch.randint(3, cfg.vocab_size, (N, 5), generator=g)], dim=1).cuda()
for cur in (1, 3, 6):
    gen.last_logits(st, ids[:, :cur].contiguous(), rd)
st_i = gen.encode(*args, beams)
for cur in range(1, 7):
    gen.step_logits(st_i, ids[:, :cur].contiguous(), rd)
del st_i
torch.cuda.empty_cache()
outs, mems, kvs, logs = [], [], [], []
kw = dict(num_beams=beams, max_length=12, no_repeat_ngram_size=3, early_stopping=True)
for call in range(3):
    mem = gen.encode_memory(*args)
    mems.append((mem.MEM.clone(), mem.mem_valid.clone(), mem.ent_valid.clone(), mem.inv_n.clone()))
    o = gen.generate_from_memory(mem, **kw)
    plan = next(iter(gen._plans.values()))
    kvs.append([t.clone() for t in plan["st"].mem.kv])
    logs.append(plan["st"].cws["logits"].clone())
    outs.append(o.clone())
    print("call", call, "shape", tuple(o.shape), flush=True)
for i in (1, 2):
    print("call 0 vs", i, "out equal:", outs[0].shape == outs[i].shape and torch.equal(outs[0], outs[i]),
          "| MEM equal:", torch.equal(mems[0][0], mems[i][0]), "valid equal:", all(torch.equal(a, b) for a, b in zip(mems[0][1:], mems[i][1:])),
          "| kv equal:", [torch.equal(a, b) for a, b in zip(kvs[0], kvs[i])], "| last logits equal:", torch.equal(logs[0], logs[i]))
    if not torch.equal(mems[0][0], mems[i][0]):
        d = (mems[0][0] != mems[i][0]).any(1).nonzero().flatten()
        print("   MEM rows differing:", d.numel(), d[:10].tolist(), "nan rows in call i:", int(torch.isnan(mems[i][0].float()).any(1).sum()))
