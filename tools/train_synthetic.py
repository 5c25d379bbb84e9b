#!/usr/bin/env python
"""`main()` of src/multimodal_train.py:441-485 on the drop-in, with synthetic businesses instead of the Yelp / Amazon files
(there is no dataset, tokenizer or pretrained checkpoint in this image): same flags, same order of statements.

  python tools/train_synthetic.py --dataset yelp --batch_size 16 --num_epochs 1 --n_train 64 --n_val 16
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_synthetic.py --batch_size 16

Every piece is the package's public surface: MultimodalSum / TableEncoder (modules.py), get_optimizer / get_scheduler (fused clip +
AdamW, optim.py), DistributedDataParallel (dp.py), SyntheticDataset (synth.py), set_environments / make_loops / train_model
(train_utils.py: the prefetchers, the epoch bodies and the checkpoint files).  Not part of the measured path (bench.py is).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def str2bool(v):                                            # src/utils.py:14-22
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


def parse(argv=None):
    p = argparse.ArgumentParser()                           # the flags of src/multimodal_train.py:411-439 ...
    p.add_argument("--workers", type=int, default=4)
    p.add_argument("--local_rank", type=int, default=0)
    p.add_argument("--dataset", type=str, default="yelp")
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--num_epochs", type=int, default=5)
    p.add_argument("--warmup_ratio", type=float, default=0.05)
    p.add_argument("--max_grad_norm", type=int, default=1)
    p.add_argument("--learning_rate", type=float, default=1e-5)
    p.add_argument("--label_smoothing", type=float, default=0.1)
    p.add_argument("--early_stopping", type=str2bool, default=False)
    p.add_argument("--bart_pretrained", type=str, default=None)
    p.add_argument("--table_pretrained", type=str, default=None)
    p.add_argument("--img_pretrained", type=str, default=None)
    p.add_argument("--n_train", type=int, default=64)       # ... plus the size of the synthetic corpus
    p.add_argument("--n_val", type=int, default=16)
    p.add_argument("--ckpt", type=str, default=None)
    return p.parse_args(argv)


def main(argv=None):
    import torch
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler
    from multimodalsum_b200.dp import DistributedDataParallel as DDP
    from multimodalsum_b200.modules import AmazonTableEncoder, MultimodalSum, YelpTableEncoder
    from multimodalsum_b200.synth import ModelConfig, SyntheticDataset
    from multimodalsum_b200.train_utils import get_optimizer, get_scheduler, make_loops, set_environments, train_model

    args = parse(argv)
    args.ckpt = args.ckpt or "ckpt/multimodal_trained_%s" % args.dataset
    args = set_environments(args)
    torch.cuda.set_device(args.local_rank)
    TableEncoder = {"yelp": YelpTableEncoder, "amazon": AmazonTableEncoder}[args.dataset]

    # Model (:459-460); random init by the reference's recipe unless checkpoint directories are given
    torch.manual_seed(0)
    cfg = ModelConfig(dataset=args.dataset)
    model = MultimodalSum(args.bart_pretrained, args.table_pretrained, args.img_pretrained, TableEncoder, config=cfg,
                          label_smoothing=args.label_smoothing)
    model.cuda()

    # Optimizer (:463-464): the fused clip + AdamW over the engine's arenas; `named_parameters()` is a generator, so the
    # no-decay group stays empty exactly as in the reference (quirk Q1)
    no_decay = ["bias", "bn1.weight", "bn2.weight", "bn3.weight", "layer_norm.weight", "layernorm_embedding.weight"]
    engine = model._ensure_engine(torch.device("cuda", args.local_rank))
    optimizer = get_optimizer(engine, args.learning_rate, no_decay, model.named_parameters(), None, max_grad_norm=args.max_grad_norm)

    # Dataset (:467, src/train_utils.py:33-47)
    amazon = args.dataset == "amazon"
    kw = dict(len_range=(45, 70) if amazon else (60, 100))
    data_train = SyntheticDataset(cfg, args.n_train, seed=1000, **kw)
    data_val = SyntheticDataset(cfg, args.n_val, seed=2000, **kw)
    train_sampler = DistributedSampler(data_train, shuffle=True) if args.distributed else None
    val_sampler = DistributedSampler(data_val, shuffle=False) if args.distributed else None
    train_dataloader = DataLoader(data_train, args.batch_size, shuffle=(train_sampler is None), num_workers=args.workers,
                                  pin_memory=True, sampler=train_sampler, drop_last=True)
    val_dataloader = DataLoader(data_val, args.batch_size, shuffle=False, num_workers=args.workers, pin_memory=True,
                                sampler=val_sampler, drop_last=False)

    field = data_train.field.cuda()                         # :469-470
    if args.distributed:
        model = DDP(model, delay_allreduce=True)            # :473-474

    t_epoch = len(train_dataloader)                         # :477-479
    args.log_interval = max(1, int(t_epoch * args.warmup_ratio))
    scheduler = get_scheduler(args, t_epoch, optimizer)

    train, validate = make_loops(args, field)
    return train_model(args, model, train_sampler, train_dataloader, val_dataloader, train, validate, optimizer, scheduler, t_epoch, "whole")


if __name__ == "__main__":
    main()
