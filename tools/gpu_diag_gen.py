import sys, os, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from test_generate_gpu import _setup
from oracle import mmsum_oracle as OR
name = sys.argv[1] if len(sys.argv) > 1 else "gen_small_yelp_s128"
gen, cfg, sd, batch, gk, ref = _setup(name)
p = {k: v.cuda() for k, v in sd.items()}
torch.backends.cuda.matmul.allow_tf32 = False
st = gen.encode(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask, 1)
# memory parity
text, tv, table, tabv, img, iv = OR.multimodal_memories(p, cfg, batch)
B, R, S = batch.reviews.shape
Sp, T, F = st["Sp"], st["T"], st["F"]
MEMtext = None
# recompute MEM from kv is not possible; re-run encode pieces: use engine workspace? we only have kv. compare kv of layer 0 instead
c = "bart_model.model.decoder.layers.0.encoder_attn."
memo = torch.cat([torch.nn.functional.pad(text, (0, 0, 0, Sp - S)).reshape(-1, 1024), table.reshape(-1, 1024), img.reshape(-1, 1024)])
kvo = torch.cat([torch.nn.functional.linear(memo, p[c + "k_proj.weight"], p[c + "k_proj.bias"]), torch.nn.functional.linear(memo, p[c + "v_proj.weight"], p[c + "v_proj.bias"])], 1)
kvc = st["kv"][0].float()
valid_rows = st["mem_valid"].bool()
d = (kvc - kvo)[valid_rows]
print("kv layer0 (valid rows): max abs err %.4f, ref abs max %.3f, rel fro %.4f" % (d.abs().max().item(), kvo[valid_rows].abs().max().item(), d.norm().item() / kvo[valid_rows].norm().item()))
for nm, lo, hi in (("text", 0, T), ("table", T, T + B * F), ("img", T + B * F, kvc.shape[0])):
    m = valid_rows[lo:hi]
    dd = (kvc[lo:hi] - kvo[lo:hi])[m]
    print("  ", nm, "rel fro %.4f" % (dd.norm().item() / max(kvo[lo:hi][m].norm().item(), 1e-9)), "valid rows", int(m.sum()))
ofn = OR.generation_logits_fn(p, cfg, batch, 1)
rd = torch.zeros(B, device="cuda")
ref = ref.cuda()
for cur in range(1, min(ref.shape[1], 8)):
    ids = ref[:, :cur].contiguous()
    lc = torch.log_softmax(gen.last_logits(st, ids, rd).float(), -1)
    lo_ = torch.log_softmax(ofn(ids), -1)
    print("step", cur, "max |dlogp| %.4f" % (lc - lo_).abs().max().item(), "logit std %.3f" % ofn(ids).std().item(), "argmax", lc.argmax(-1).tolist(), lo_.argmax(-1).tolist())
print("per-business text kv rel err:")
for b in range(B):
    lo, hi = b * R * Sp, (b + 1) * R * Sp
    m = valid_rows[lo:hi]
    dd = (kvc[lo:hi] - kvo[lo:hi])[m]
    print("  biz", b, "%.4f" % (dd.norm().item() / kvo[lo:hi][m].norm().item()), "valid", int(m.sum()))
# decoder internals at step 1: compare hidden states after each stage for cur=1
ids = ref[:, :1].contiguous()
_ = gen.last_logits(st, ids, rd)
w = st["ws"]
import torch.nn.functional as Fn
pre = "bart_model.model.decoder."
x0 = OR.embed(ids, p, pre, torch.zeros(B, 1, device="cuda"))
N = B
xc = w["x"] if cfg.decoder_layers % 2 == 0 else w["nxt"]
# recompute the oracle decoder to get the final hidden state of position 0
mems = [text, table, img]; valids = [tv, tabv, iv]
xo = OR.decoder(p, cfg, ids, mems, valids, torch.zeros(B, 1, device="cuda"), use_pad_mask=False)
xf = xc.view(N, 128, 1024)[:, 0].float()
for b in range(B):
    print("final hidden biz", b, "rel err %.4f" % ((xf[b] - xo[b, 0]).norm().item() / xo[b, 0].norm().item()))
# layer-0 pieces: x1 after self-attn block (position 0 attends only to itself)
print("pres", st["pres"].tolist(), "inv_n", st["inv_n"].tolist(), "ent_valid", st["ent_valid"].tolist())
