import sys, os
sys.path.insert(0, ".")
import torch
from multimodalsum_b200.synth import ModelConfig, make_batch, make_state_dict
from multimodalsum_b200.generation import Generator
from multimodalsum_b200.modules import MultimodalSum, YelpTableEncoder
from oracle import mmsum_oracle as OR
torch.backends.cuda.matmul.allow_tf32 = False
for std, go, S in ((0.02, False, 128), (0.02, True, 128), (0.08, True, 128), (0.02, True, 150)):
    cfg = ModelConfig(encoder_layers=2, decoder_layers=2, ffn_dim=256, vocab_size=512, max_position_embeddings=256, dropout=0.0, dataset="yelp", init_std=std)
    sd = make_state_dict(cfg, seed=21, gates_open=go)
    batch = make_batch(cfg, 3, seed=31, n_reviews=3, max_imgs=2, seq_len=S, len_range=(S - 60, S - 10)).to("cuda")
    model = MultimodalSum(TableEncoder=YelpTableEncoder, config=cfg); model.load_state_dict(sd, strict=False); model = model.cuda().eval()
    gen = Generator(model)
    p = {k: v.cuda() for k, v in sd.items()}
    B = 3
    st = gen.encode(batch.reviews, batch.reviews_mask, batch.field, batch.field_value, batch.img, batch.img_mask, 1)
    ofn = OR.generation_logits_fn(p, cfg, batch, 1)
    g = torch.Generator(device="cuda").manual_seed(5)
    ids = torch.randint(3, 512, (B, 9), device="cuda", generator=g); ids[:, 0] = 2
    rd = torch.zeros(B, device="cuda")
    errs = []
    for cur in (1, 4, 9):
        lc = torch.log_softmax(gen.last_logits(st, ids[:, :cur].contiguous(), rd).float(), -1)
        lo = torch.log_softmax(ofn(ids[:, :cur].contiguous()), -1)
        errs.append([round((lc[b] - lo[b]).abs().max().item(), 4) for b in range(B)])
    print("std", std, "gates_open", go, "S", S, "img_mask", batch.img_mask.tolist(), "max|dlogp| per business at cur=1,4,9:", errs)
