"""Train-step glue of the reference around the fused step.

  get_optimizer / get_scheduler       src/train_utils.py:49-63       (re-exported from .optim, incl. quirk Q1)
  train_step                          src/multimodal_train.py:357-364 forward, zero_grad, backward (+ bucketed all-reduce),
                                      clip_grad_norm_, AdamW.step, scheduler.step
  save_checkpoint                     src/train_utils.py:79-97       pytorch_model.bin (+ training_state.bin), `save_option`
                                      'whole' | 'text' | 'img' | 'table' selects the sub-module whose state_dict is written
  load_checkpoint                     (no reference counterpart)     resume: weights, optimizer moments, schedule position
  set_environments                    src/train_utils.py:12-31       checkpoint dir + training_args.bin, NCCL process group
  make_loops -> (train, validate)     src/multimodal_train.py:346-408 the epoch bodies with the reference's signatures (the
                                      module globals `args` / `field` they read are bound here), on prefetch.*_data_prefetcher;
                                      `stage=` 'text' | 'img' | 'table' gives the loops of src/text_pretrain.py:151-207,
                                      src/img_pretrain.py:178-238, src/table_pretrain.py:242-303
  make_test_loop -> test              src/test.py:137-165            generation over a loader (get_multimodal_outputs + generate)
  train_model                         src/train_utils.py:65-97       epochs, sampler epochs, validation, early stopping, files
  AverageMeter                        src/utils.py:40-56
The files interchange with the reference: `state_dict` keys / shapes are the reference's, the optimizer state
uses transformers-AdamW names.
"""
import datetime
import os
import time

import torch

from .dp import reduce_tensor
from .optim import FusedAdamW, LinearWarmupSchedule, get_optimizer, get_scheduler  # noqa: F401


def train_step(model, optimizer, scheduler, inputs, reducer=None):
    """One optimizer step exactly in the reference's order.  `optimizer` is a FusedAdamW built with `max_grad_norm`
    (the clip is folded into the update) or any torch optimizer (then `clip_grad_norm_` runs as in the reference)."""
    loss = model(*inputs)[0]
    optimizer.zero_grad()
    loss.backward()                      # a GradAllReducer (dp.py) hooks the engine and overlaps the all-reduce with this call
    if not isinstance(optimizer, FusedAdamW):
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss


def save_checkpoint(model, optimizer, scheduler, epoch, ckpt_dir, save_option="whole"):
    """src/train_utils.py:79-97 (rank 0 only in the reference): writes `pytorch_model.bin` and `training_state.bin`."""
    os.makedirs(ckpt_dir, exist_ok=True)
    save_model = getattr(model, "module", model)
    if save_option == "text":
        save_model = save_model.bart_model
    elif save_option == "img":
        save_model = save_model.img_encoder
    elif save_option == "table":
        save_model = save_model.table_encoder
    sd = {k: v.detach().cpu().clone() for k, v in save_model.state_dict().items()}
    torch.save(sd, os.path.join(ckpt_dir, "pytorch_model.bin"))
    state = {"epoch": epoch, "optimizer": optimizer.state_dict() if optimizer is not None else None,
             "scheduler": scheduler.state_dict() if scheduler is not None else None}
    eng = getattr(getattr(model, "module", model), "engine", None)
    if eng is not None:
        state["engine"] = eng.rng_state()        # dropout seed + step counter (a key the reference never reads)
    torch.save(state, os.path.join(ckpt_dir, "training_state.bin"))
    return ckpt_dir


def load_checkpoint(model, optimizer, scheduler, ckpt_dir, save_option="whole", map_location="cpu"):
    """Resume from the files `save_checkpoint` (or the reference, src/train_utils.py:96-97) wrote — the reference itself never
    reads `training_state.bin` back, so this is an addition: weights into the (sub-)module `save_option` names,
    optimizer moments / step and the schedule position restored, the next epoch index returned.  The engine's bf16 compute copy
    follows at the next forward (parameter versions change)."""
    target = getattr(model, "module", model)
    if save_option == "text":
        target = target.bart_model
    elif save_option == "img":
        target = target.img_encoder
    elif save_option == "table":
        target = target.table_encoder
    sd = torch.load(os.path.join(ckpt_dir, "pytorch_model.bin"), map_location=map_location)
    missing = target.load_state_dict(sd, strict=False)      # missing keys are tolerated as in the stage hand-offs (:116-122)
    if missing.unexpected_keys:
        raise KeyError("unexpected keys in %s/pytorch_model.bin: %s" % (ckpt_dir, missing.unexpected_keys[:4]))
    eng = getattr(getattr(model, "module", model), "engine", None)
    if eng is not None:
        eng.mark_weights_dirty()
    state = torch.load(os.path.join(ckpt_dir, "training_state.bin"), map_location=map_location)
    if optimizer is not None and state.get("optimizer") is not None:
        optimizer.load_state_dict(state["optimizer"])
    if scheduler is not None and state.get("scheduler") is not None:
        scheduler.load_state_dict(state["scheduler"])
    if eng is not None and state.get("engine") is not None:
        eng.set_rng_state(state["engine"])       # the resumed run draws the dropout masks the uninterrupted run would have
    return int(state["epoch"]) + 1


class AverageMeter:
    """src/utils.py:40-56."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def set_environments(args):
    """src/train_utils.py:12-31: rank 0 creates `args.ckpt` and writes training_args.bin; under torchrun (WORLD_SIZE > 1) the
    process binds its GPU and joins the NCCL group (one process per GPU).  LOCAL_RANK from the environment wins over
    `args.local_rank` (torchrun no longer passes --local_rank)."""
    args.local_rank = int(os.environ.get("LOCAL_RANK", getattr(args, "local_rank", 0)))
    if args.local_rank == 0:
        os.makedirs(args.ckpt, exist_ok=True)
        torch.save(vars(args), "%s/training_args.bin" % args.ckpt)
    args.distributed = int(os.environ.get("WORLD_SIZE", "1")) > 1
    args.gpu = 0
    args.world_size = 1
    if args.distributed:
        args.gpu = args.local_rank
        torch.cuda.set_device(args.gpu)
        if not torch.distributed.is_initialized():
            torch.distributed.init_process_group(backend="nccl", init_method="env://", device_id=torch.device("cuda", args.gpu))
        args.world_size = torch.distributed.get_world_size()
        print("GPU {}/{}".format(args.gpu, args.world_size))
    return args


def _stage_table(args, field, stage):
    """Per training stage: the prefetcher, how a staged tuple becomes the model call, the batch size `validate` weights by, and
    the parameters `clip_grad_norm_` sees (all for the multimodal / text stages; the trained head only for the img / table
    pretraining stages, src/img_pretrain.py:190-194, src/table_pretrain.py:259-263)."""
    from . import prefetch
    by_dataset = lambda y, a: {"yelp": y, "amazon": a}.get(args.dataset)       # noqa: E731
    unwrap = lambda m: getattr(m, "module", m)                                  # noqa: E731
    if stage == "multimodal":        # src/multimodal_train.py:346-408
        return dict(prefetcher=by_dataset(prefetch.yelp_data_prefetcher, prefetch.amazon_data_prefetcher),
                    call=lambda m, t: m(t[0], t[1], t[2], field, t[3], t[4], t[5]), first=lambda t: t[0],
                    clip=lambda m: m.parameters())
    if stage == "text":              # src/text_pretrain.py:151-207
        return dict(prefetcher=prefetch.text_data_prefetcher, call=lambda m, t: m(t[0], t[1], t[2]), first=lambda t: t[0],
                    clip=lambda m: m.parameters())
    if stage == "img":               # src/img_pretrain.py:178-238
        return dict(prefetcher=prefetch.img_data_prefetcher, call=lambda m, t: m(t[0], t[1], labels=t[2]), first=lambda t: t[0],
                    clip=lambda m: [p for n, p in unwrap(m).named_parameters() if n.startswith("img_encoder")])
    if stage == "table":             # src/table_pretrain.py:242-303
        return dict(prefetcher=by_dataset(prefetch.yelp_table_data_prefetcher, prefetch.amazon_table_data_prefetcher),
                    call=lambda m, t: m(field, t[0], labels=t[1]), first=lambda t: t[0][0],
                    clip=lambda m: [p for n, p in unwrap(m).table_encoder.named_parameters() if not n.startswith("bart")])
    raise ValueError("stage must be 'multimodal', 'text', 'img' or 'table'")


def make_loops(args, field=None, log=print, stage="multimodal", device=None):
    """-> (train, validate) with the signatures of src/multimodal_train.py:346 / :381 (and of the three pretraining scripts'
    loops, `stage=`), to be handed to `train_model`.

    `args` needs: dataset ('yelp' | 'amazon'), max_grad_norm, log_interval, distributed, world_size, local_rank.  `field` is the
    table's field-name ids already on the GPU (src/multimodal_train.py:469-470; unused by the text / img stages).  Statement
    order is the reference's; batches come through the staging prefetchers (prefetch.py); with a FusedAdamW built with
    `max_grad_norm` the clip is part of the update kernel, otherwise `clip_grad_norm_` runs as in the reference."""
    st = _stage_table(args, field, stage)
    if st["prefetcher"] is None:
        raise ValueError("args.dataset must be 'yelp' or 'amazon'")
    on_cuda = device is None or torch.device(device).type == "cuda"

    def _sync():
        if on_cuda:
            torch.cuda.synchronize()

    def train(start_time, train_dataloader, model, optimizer, scheduler, e, t_epoch):
        model.train()
        prefetcher = st["prefetcher"](train_dataloader, device=device)
        batch = prefetcher.next()
        i = 0
        fused_clip = isinstance(optimizer, FusedAdamW) and optimizer.max_grad_norm is not None
        while st["first"](batch) is not None:
            loss = st["call"](model, batch)[0]

            optimizer.zero_grad()
            loss.backward()
            if args.max_grad_norm is not None and not fused_clip:
                torch.nn.utils.clip_grad_norm_(st["clip"](model), args.max_grad_norm)
            optimizer.step()
            scheduler.step()

            if args.log_interval and i % args.log_interval == 0:
                reduced_loss = reduce_tensor(loss.data, args.world_size) if args.distributed else loss.data
                _sync()
                if args.local_rank == 0:
                    timedelta = str(datetime.timedelta(seconds=int(time.time() - start_time)))
                    log("{} epoch {} batch id {}/{} loss {}".format(timedelta, e + 1, i + 1, t_epoch, reduced_loss.item()))

            batch = prefetcher.next()
            i += 1
        return i

    def validate(val_dataloader, model, e):
        model.eval()
        losses = AverageMeter()
        prefetcher = st["prefetcher"](val_dataloader, device=device)
        batch = prefetcher.next()
        while st["first"](batch) is not None:
            with torch.no_grad():
                loss = st["call"](model, batch)[0]
            reduced_loss = reduce_tensor(loss.data, args.world_size) if args.distributed else loss.data
            losses.update(reduced_loss.item(), st["first"](batch).size(0))
            batch = prefetcher.next()
        _sync()
        if args.local_rank == 0:
            log("{} epoch valid loss {}".format(e + 1, losses.avg))
        return losses.avg

    return train, validate


def make_test_loop(args, field, log=print, device=None):
    """-> `test(test_dataloader, model, tokenizer)` of src/test.py:137-165: beam-search summaries for every business of the loader.
    `args` needs dataset, num_beams, length_penalty, max_length.  The body is the reference's: memories from
    `model.get_multimodal_outputs`, zero `rating_diff`, `model.bart_model.generate(..., no_repeat_ngram_size=3,
    early_stopping=True)`; `tokenizer` may be None (token-id lists are returned undecoded)."""
    from . import prefetch
    Prefetcher = {"yelp": prefetch.yelp_data_prefetcher, "amazon": prefetch.amazon_data_prefetcher}.get(args.dataset)
    if Prefetcher is None:
        raise ValueError("args.dataset must be 'yelp' or 'amazon'")

    def test(test_dataloader, model, tokenizer):
        model.eval()
        prefetcher = Prefetcher(test_dataloader, device=device)
        reviews, reviews_mask, reviews_rating, field_value, img, img_mask = prefetcher.next()
        i = 0
        generated_list = []
        while reviews is not None:
            try:
                log("%d / %d" % (i + 1, len(test_dataloader)))
            except TypeError:                      # a loader without a length
                log("%d" % (i + 1))
            with torch.no_grad():
                _, text_hiddens, text_attention_mask, table_hiddens, table_attention_mask, img_hiddens, img_attention_mask = \
                    model.get_multimodal_outputs(reviews, reviews_mask, field, field_value, img, img_mask)
                rating_diff = torch.zeros([text_hiddens.size(0), 1], device=text_hiddens.device)
                generated = model.bart_model.generate(text_hiddens, text_attention_mask, table_hiddens, table_attention_mask,
                                                      img_hiddens, img_attention_mask, rating_diff=rating_diff,
                                                      num_beams=args.num_beams, length_penalty=args.length_penalty,
                                                      max_length=args.max_length, no_repeat_ngram_size=3, early_stopping=True)
            generated_list.extend(generated)
            reviews, reviews_mask, reviews_rating, field_value, img, img_mask = prefetcher.next()
            i += 1
        if tokenizer is not None:
            generated_list = [tokenizer.decode(g, skip_special_tokens=True, clean_up_tokenization_spaces=False) for g in generated_list]
        return generated_list

    return test


def train_model(args, model, train_sampler, train_dataloader, val_dataloader, train, validate, optimizer, scheduler, t_epoch,
                save_option="whole", log=print):
    """src/train_utils.py:65-97: per epoch — sampler / dataset epoch bookkeeping, `train`, `validate`, and on rank 0 the
    checkpoint files when the validation loss is the best so far (`args.early_stopping`) or always (otherwise).  Returns the
    list of validation losses (rank 0's view)."""
    start_time = time.time()
    val_loss_list = []
    for e in range(args.num_epochs):
        log("Epoch {}".format(e + 1))
        if args.distributed and train_sampler is not None:
            train_sampler.set_epoch(e)
        if e != 0 and hasattr(getattr(train_dataloader, "dataset", None), "set_epoch"):
            train_dataloader.dataset.set_epoch()

        train(start_time, train_dataloader, model, optimizer, scheduler, e, t_epoch)
        val_loss = validate(val_dataloader, model, e)

        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if args.local_rank == 0:
            val_loss_list.append(val_loss)
            if (not args.early_stopping) or val_loss <= min(val_loss_list):
                save_checkpoint(model, optimizer, scheduler, e, args.ckpt, save_option)
    return val_loss_list
