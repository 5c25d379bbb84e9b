"""Host -> device input pipeline of the training / evaluation loops: the reference's `*_data_prefetcher` classes
(src/multimodal_train.py:196-343, src/img_pretrain.py:144-176, src/table_pretrain.py:132-241), same constructor, same `next()`
tuples, same end-of-epoch convention (every element None), so `train()` / `validate()` (src/multimodal_train.py:344-404) run
unchanged.

What differs from the reference is how a batch gets to HBM.  The reference allocates fresh device tensors for every batch
(`.cuda(non_blocking=True)` on a side stream) and hands them to the caching allocator's cross-stream bookkeeping with
`record_stream`.  At B200 step times (≈ 90 ms for 129 MB of inputs) an occasional `cudaMalloc` / deferred free inside the loop
is visible as 5-10 % dips, so here a batch is copied into one of `n_stage` RESIDENT staging slots:

* slot tensors are allocated once per shape and re-used; `next()` hands out views of the slot (narrowed along dim 0 for the short
  last batch of a `drop_last=False` loader);
* the copy of batch i+1 is enqueued on a copy stream while the step of batch i runs; the consumer stream waits on the copy's
  event, and a slot is overwritten only after the work that was enqueued while it was the current batch has finished (an
  event recorded on the consumer stream at the following `next()` — the loop's `backward()` / `optimizer.step()` of that batch
  are enqueued before it, exactly the ordering `record_stream` relies on in the reference);
* sources that are not in pinned memory (a DataLoader without `pin_memory=True`) go through a pinned host slot first, so the
  H2D copy is asynchronous either way;
* ids / masks are normalised to the dtypes the kernels read (int64 ids and masks, fp32 ratings, bool image masks) on the HOST,
  before the copy — no conversion kernels on the device;
* the longest review of the batch is measured on the host mask and attached to the returned `reviews_mask` as
  `.max_review_len`: `MultimodalSum.forward` / `TextSupervised.forward` pick it up and trim the encoder frames to it without
  reading the mask back from the device.

Nothing here computes on the data; it is plumbing around `torch.cuda.Stream` / `torch.cuda.Event`.  With `device="cpu"` the same
slot logic runs without streams (host-logic tests only: the modules themselves refuse CPU tensors).
"""
import contextlib

import torch

__all__ = ["StagedPrefetcher", "yelp_data_prefetcher", "amazon_data_prefetcher", "text_data_prefetcher", "img_data_prefetcher",
           "yelp_table_data_prefetcher", "amazon_table_data_prefetcher"]


def _canonical(name, t):
    """Host-side dtype normalisation (what `modules._ids` / `.float()` / `.to(bool)` would otherwise do on the device)."""
    if name in ("img_mask", "input_imgs_mask"):
        return t if t.dtype == torch.bool else t.ne(0)
    if t.dtype in (torch.bfloat16, torch.float32):
        return t
    if t.is_floating_point():
        return t.float()
    if name == "reviews_rating":
        return t.float()
    return t if t.dtype == torch.int64 else t.long()


class _Slot:
    """One resident staging batch: device tensors (+ pinned host mirrors for pageable sources), the event of the last copy into it
    and the event after which it may be overwritten."""

    def __init__(self):
        self.dev = {}
        self.pinned = {}
        self.rows = {}
        self.ready = None
        self.free = None
        self.hint = None
        self.seen_streams = set()     # consumer streams these tensors are registered with (record_stream)


class StagedPrefetcher:
    """Generic ring: `names` = the loader tuple's field names in order; `groups` describes the tuple `next()` returns — a name, or a
    list of names (returned as a list, e.g. the table's `field_value`)."""

    names = ()
    groups = ()
    mask_name = "reviews_mask"

    def __init__(self, loader, device=None, n_stage=3):
        if n_stage < 2:
            raise ValueError("n_stage >= 2: one slot is consumed while the next is filled")
        self.loader = iter(loader)
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("the prefetcher stages batches in CUDA memory; no CUDA device is visible")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self.slots = [_Slot() for _ in range(n_stage)]
        self.n_staged = 0          # batches copied so far (slot of batch i = i % n_stage)
        self.n_served = 0
        self.pending = None        # slot index holding the batch that the next `next()` returns
        self.last_served = None
        self.h2d_bytes = 0
        self.allocations = 0       # device (re)allocations: 1 per field and slot for a fixed-shape loader
        self.preload()

    # ------------------------------------------------------------------ staging
    def _slot_tensor(self, slot, name, src):
        """Device tensor of `slot` for field `name`, large enough along dim 0 for `src` (re-allocated only on a shape / dtype change
        other than a shorter leading dimension).  Allocated on the caller's current stream and registered with the copy stream
        (`record_stream`), so that the caching allocator orders a later re-use of the block after both."""
        cur = slot.dev.get(name)
        if (cur is None or cur.dtype != src.dtype or cur.dim() != src.dim() or cur.shape[1:] != src.shape[1:]
                or (src.dim() > 0 and cur.shape[0] < src.shape[0])):
            cur = torch.empty(src.shape, dtype=src.dtype, device=self.device)
            if self.cuda:
                cur.record_stream(self.stream)
            slot.dev[name] = cur
            slot.pinned.pop(name, None)
            slot.seen_streams = set()
            self.allocations += 1
        return cur

    @staticmethod
    def _rows_view(t, rows):
        return t if rows is None or rows == t.shape[0] else t.narrow(0, 0, rows)

    def _stage_host(self, slot, name, src):
        """Pageable source -> the slot's pinned mirror (a host memcpy); pinned sources are copied from where they are."""
        if not self.cuda or src.is_pinned():
            return src
        dst = slot.dev[name]
        pin = slot.pinned.get(name)
        if pin is None or pin.shape != dst.shape or pin.dtype != dst.dtype:
            pin = torch.empty(dst.shape, dtype=dst.dtype, pin_memory=True)
            slot.pinned[name] = pin
        pv = self._rows_view(pin, src.shape[0] if src.dim() > 0 else None)
        pv.copy_(src)
        return pv

    def preload(self):
        """Pull the next batch from the loader and enqueue its copy into the next slot (on the copy stream)."""
        try:
            items = next(self.loader)
        except StopIteration:
            self.pending = None
            return
        if len(items) != len(self.names):
            raise ValueError("%s expects %d fields per batch (%s), the loader produced %d"
                             % (type(self).__name__, len(self.names), ", ".join(self.names), len(items)))
        k = self.n_staged % len(self.slots)
        slot = self.slots[k]
        host = {}
        for name, t in zip(self.names, items):
            if not torch.is_tensor(t):
                raise TypeError("field %r: expected a tensor, got %s" % (name, type(t).__name__))
            if t.device.type != "cpu":
                raise ValueError("field %r already lives on %s: the prefetcher stages HOST batches" % (name, t.device))
            host[name] = _canonical(name, t).contiguous()
        slot.hint = None
        if self.mask_name in host:
            m = host[self.mask_name]
            S = m.shape[-1]
            slot.hint = int((m.ne(0).to(torch.int32) * torch.arange(1, S + 1, dtype=torch.int32)).max().item()) if m.numel() else 0
        if self.cuda and slot.ready is not None and slot.pinned:
            slot.ready.synchronize()            # the pinned mirrors are rewritten by the host below: their last H2D copy is done
        n_alloc = self.allocations
        for name, t in host.items():            # (re)allocation happens on the caller's stream, outside the copy-stream scope
            self._slot_tensor(slot, name, t)
            slot.rows[name] = t.shape[0] if t.dim() > 0 else None
            host[name] = self._stage_host(slot, name, t)
        if self.cuda and self.allocations != n_alloc:
            # a new block may be recycled memory whose last use is still queued on the caller's stream (the caching allocator
            # orders re-use only within the allocating stream): the copy stream must not write into it before that work is done.
            # Happens for the first n_stage batches (and on a shape change), never in steady state.
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with (torch.cuda.stream(self.stream) if self.cuda else contextlib.nullcontext()):
            if self.cuda and slot.free is not None:
                self.stream.wait_event(slot.free)
            for name, t in host.items():
                self._rows_view(slot.dev[name], slot.rows[name]).copy_(t, non_blocking=True)
                self.h2d_bytes += t.numel() * t.element_size()
            if self.cuda:
                if slot.ready is None:
                    slot.ready = torch.cuda.Event()
                slot.ready.record(self.stream)
        self.pending = k
        self.n_staged += 1

    # ------------------------------------------------------------------ consumer side
    def _view(self, slot, name):
        t = slot.dev[name]
        rows = slot.rows[name]
        v = t.narrow(0, 0, rows) if rows is not None and rows != t.shape[0] else t.view(t.shape)   # a fresh tensor object per batch
        if name == self.mask_name and slot.hint is not None:
            v.max_review_len = slot.hint
        return v

    def next(self):
        """The staged batch as the reference's tuple (views of a slot), or all None at the end of the loader; the following batch
        starts copying before this returns.  Contract (the reference loop's own pattern): all GPU work that reads a batch is
        enqueued before `next()` is called again — the slot is re-used `n_stage - 1` calls later and its rewrite waits only for
        work enqueued up to the call that followed its own."""
        cur = torch.cuda.current_stream(self.device) if self.cuda else None
        if self.cuda and self.last_served is not None:
            # everything enqueued so far used the previous batch at the latest: its slot is free once that work is done
            prev = self.slots[self.last_served]
            if prev.free is None:
                prev.free = torch.cuda.Event()
            prev.free.record(cur)                 # (a wait already enqueued on the copy stream keeps the earlier record)
        if self.pending is None:
            self.last_served = None
            return self._pack(None)
        k = self.pending
        slot = self.slots[k]
        if self.cuda:
            cur.wait_event(slot.ready)
            if cur.cuda_stream not in slot.seen_streams:       # once per (slot, consumer stream): see _slot_tensor
                for t in slot.dev.values():
                    t.record_stream(cur)
                slot.seen_streams.add(cur.cuda_stream)
        out = self._pack(slot)
        self.last_served = k
        self.n_served += 1
        self.preload()
        return out

    def _pack(self, slot):
        out = []
        for g in self.groups:
            if isinstance(g, (list, tuple)):
                out.append([None if slot is None else self._view(slot, n) for n in g])
            else:
                out.append(None if slot is None else self._view(slot, g))
        return tuple(out)

    # iteration sugar (not in the reference): `for batch in prefetcher:`
    def __iter__(self):
        return self

    def __next__(self):
        out = self.next()
        first = out[0][0] if isinstance(out[0], list) else out[0]
        if first is None:
            raise StopIteration
        return out


_YELP_TABLE = ["name", "category", "str_categorical", "str_boolean", "rating", "hours"]
_AMAZON_TABLE = ["price", "rating", "brand", "name", "category", "description"]


class yelp_data_prefetcher(StagedPrefetcher):
    """src/multimodal_train.py:196-268 -> (reviews, reviews_mask, reviews_rating, [name, category, str_categorical, str_boolean,
    rating, hours], img, img_mask)."""
    names = ("reviews", "reviews_mask", "reviews_rating", *_YELP_TABLE, "img", "img_mask")
    groups = ("reviews", "reviews_mask", "reviews_rating", _YELP_TABLE, "img", "img_mask")


class amazon_data_prefetcher(StagedPrefetcher):
    """src/multimodal_train.py:271-343 -> (reviews, reviews_mask, reviews_rating, [price, rating, brand, name, category,
    description], img, img_mask)."""
    names = ("reviews", "reviews_mask", "reviews_rating", *_AMAZON_TABLE, "img", "img_mask")
    groups = ("reviews", "reviews_mask", "reviews_rating", _AMAZON_TABLE, "img", "img_mask")


class text_data_prefetcher(StagedPrefetcher):
    """src/text_pretrain.py (`data_prefetcher`) -> (reviews, reviews_mask, reviews_rating)."""
    names = groups = ("reviews", "reviews_mask", "reviews_rating")


class img_data_prefetcher(StagedPrefetcher):
    """src/img_pretrain.py:144-176 (`data_prefetcher`) -> (input_imgs, input_imgs_mask, labels)."""
    names = groups = ("input_imgs", "input_imgs_mask", "labels")


class yelp_table_data_prefetcher(StagedPrefetcher):
    """src/table_pretrain.py:132-185 (`yelp_data_prefetcher`) -> ([name, ..., hours], label)."""
    names = (*_YELP_TABLE, "label")
    groups = (_YELP_TABLE, "label")


class amazon_table_data_prefetcher(StagedPrefetcher):
    """src/table_pretrain.py:187-241 (`amazon_data_prefetcher`) -> ([price, ..., description], label)."""
    names = (*_AMAZON_TABLE, "label")
    groups = (_AMAZON_TABLE, "label")
