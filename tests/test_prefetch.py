"""Input pipeline (multimodalsum_b200/prefetch.py) against the behaviour of the reference's `*_data_prefetcher` classes
(src/multimodal_train.py:196-343, src/img_pretrain.py:144-176, src/table_pretrain.py:132-241): same tuples, same end-of-loader
convention, contents bit-exact (byte / integer work), plus what the staging ring adds (resident slots, short last batch,
host-side dtype normalisation, the length hint).  CPU tests run the slot logic with device="cpu"; the GPU tests run the real
streams / events and the reference's training-loop body through the prefetcher."""
import pytest
import torch

from multimodalsum_b200 import prefetch as PF
from multimodalsum_b200.synth import ModelConfig, make_batch, make_state_dict

SMALL = dict(encoder_layers=1, decoder_layers=1, ffn_dim=128, vocab_size=300, max_position_embeddings=128, dropout=0.0)


def _yelp_tuple(b):
    return (b.reviews, b.reviews_mask, b.reviews_rating, *b.field_value, b.img, b.img_mask)


def _loader(cfg, sizes, seed0=10, **kw):
    out = []
    for i, n in enumerate(sizes):
        lo = 20 + 7 * i
        out.append(make_batch(cfg, n, seed=seed0 + i, n_reviews=3, max_imgs=2, len_range=(lo, lo + 9), **kw))
    return out


def _snapshot(out):
    def cp(t):
        c = t.clone()
        if hasattr(t, "max_review_len"):
            c.max_review_len = t.max_review_len
        return c
    return tuple([cp(x) for x in g] if isinstance(g, list) else cp(g) for g in out)


def _check_stream(pf_cls, batches, to_tuple, device, transform=lambda t: t, n_stage=3):
    pf = pf_cls([tuple(transform(t) for t in to_tuple(b)) for b in batches], device=device, n_stage=n_stage)
    served = []
    out = pf.next()
    while (out[0][0] if isinstance(out[0], list) else out[0]) is not None:
        served.append(_snapshot(out))                            # a slot is re-used n_stage - 1 calls later: copy the batch out
        out = pf.next()
    assert all(x is None for g in out for x in (g if isinstance(g, list) else [g]))       # end of loader: every element None
    assert len(served) == len(batches)
    return pf, served


def _flat(out):
    return [x for g in out for x in (g if isinstance(g, list) else [g])]


@pytest.mark.parametrize("dataset", ["yelp", "amazon"])
def test_multimodal_prefetcher_tuples_and_contents_cpu(dataset):
    cfg = ModelConfig(dataset=dataset, **SMALL)
    batches = _loader(cfg, [4, 4, 4, 4, 2])                     # drop_last=False: a short last batch
    cls = PF.yelp_data_prefetcher if dataset == "yelp" else PF.amazon_data_prefetcher
    pf, served = _check_stream(cls, batches, _yelp_tuple, "cpu")
    for b, out in zip(batches, served):
        reviews, reviews_mask, reviews_rating, field_value, img, img_mask = out      # the unpacking of src/multimodal_train.py:353
        assert isinstance(field_value, list) and len(field_value) == 6
        for got, want in zip(_flat(out), _yelp_tuple(b)):
            assert got.shape == want.shape and got.dtype == want.dtype
            assert torch.equal(got, want)
        assert reviews_mask.max_review_len == int(b.reviews_mask.sum(-1).max())       # make_batch masks are prefixes
    # resident slots: one allocation per field and slot, none for the later batches or the short one
    assert pf.allocations == len(cls.names) * 3
    assert pf.h2d_bytes == sum(b.nbytes() - b.field.numel() * 8 for b in batches)


def test_slot_reuse_distance():
    cfg = ModelConfig(dataset="yelp", **SMALL)
    batches = _loader(cfg, [2] * 6)
    pf = PF.yelp_data_prefetcher([_yelp_tuple(b) for b in batches], device="cpu", n_stage=3)
    first = pf.next()
    keep = [t.clone() for t in _flat(first)]
    second = pf.next()                                           # stages batch 2 into the third slot
    assert all(torch.equal(a, b) for a, b in zip(_flat(first), keep))
    assert first[0].data_ptr() != second[0].data_ptr()
    pf.next()                                                    # stages batch 3 into the first slot again
    assert torch.equal(first[0], batches[3].reviews)             # documented: the slot is re-used n_stage - 1 calls later
    # a fresh tensor object per batch carries the hint (the slot tensor itself is never annotated)
    assert not hasattr(pf.slots[0].dev["reviews_mask"], "max_review_len")


def test_host_side_dtype_normalisation_and_errors():
    cfg = ModelConfig(dataset="yelp", **SMALL)
    b = _loader(cfg, [2])[0]
    odd = (b.reviews.int(), b.reviews_mask.bool(), b.reviews_rating.double(), *[v.int() for v in b.field_value],
           b.img.double(), b.img_mask.to(torch.uint8))
    pf = PF.yelp_data_prefetcher([odd], device="cpu")
    reviews, reviews_mask, reviews_rating, field_value, img, img_mask = pf.next()
    assert reviews.dtype == torch.int64 and reviews_mask.dtype == torch.int64 and reviews_rating.dtype == torch.float32
    assert all(v.dtype == torch.int64 for v in field_value) and img.dtype == torch.float32 and img_mask.dtype == torch.bool
    assert torch.equal(reviews, b.reviews) and torch.equal(reviews_mask, b.reviews_mask) and torch.equal(img_mask, b.img_mask)
    assert torch.equal(img, b.img) and reviews_mask.max_review_len == int(b.reviews_mask.sum(-1).max())
    bf = PF.yelp_data_prefetcher([_yelp_tuple(b)[:-2] + (b.img.bfloat16(), b.img_mask)], device="cpu").next()
    assert bf[4].dtype == torch.bfloat16                         # bf16 features stay bf16 (modules.MultimodalSum.forward accepts both)
    with pytest.raises(ValueError, match="expects 11 fields"):
        PF.yelp_data_prefetcher([_yelp_tuple(b)[:-1]], device="cpu")
    with pytest.raises(TypeError, match="expected a tensor"):
        PF.yelp_data_prefetcher([_yelp_tuple(b)[:-1] + ([1, 0],)], device="cpu")
    with pytest.raises(ValueError, match="n_stage"):
        PF.yelp_data_prefetcher([_yelp_tuple(b)], device="cpu", n_stage=1)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CUDA device"):
            PF.yelp_data_prefetcher([_yelp_tuple(b)])            # default device is CUDA: no silent CPU pipeline


def test_shape_change_reallocates_only_the_changed_field():
    cfg = ModelConfig(dataset="yelp", **SMALL)
    a, c = _loader(cfg, [2, 2])
    longer = make_batch(cfg, 3, seed=77, n_reviews=3, max_imgs=2)          # a LARGER batch than the slot holds
    pf = PF.yelp_data_prefetcher([_yelp_tuple(a), _yelp_tuple(c), _yelp_tuple(a), _yelp_tuple(longer)], device="cpu", n_stage=2)
    outs = [pf.next() for _ in range(4)]
    assert pf.next()[0] is None
    assert torch.equal(outs[3][0], longer.reviews) and outs[3][0].shape[0] == 3
    assert pf.allocations == 11 * 2 + 11                         # two slots, then slot 1 regrown once for the larger batch


def test_stage_prefetchers_cpu():
    g = torch.Generator().manual_seed(0)
    img_batches = [(torch.rand(2, 4, 196, 8, generator=g), torch.rand(2, 4, generator=g) > 0.3, torch.randint(0, 50, (2, 16), generator=g))
                   for _ in range(3)]
    pf, served = _check_stream(PF.img_data_prefetcher, img_batches, lambda b: b, "cpu")
    for want, (input_imgs, input_imgs_mask, labels) in zip(img_batches, served):        # src/img_pretrain.py:182
        assert torch.equal(input_imgs, want[0]) and torch.equal(input_imgs_mask, want[1]) and torch.equal(labels, want[2])
    cfg = ModelConfig(dataset="amazon", **SMALL)
    tb = [tuple(make_batch(cfg, 2, seed=s, n_reviews=2).field_value) + (torch.randint(0, 50, (2, 16), generator=g),) for s in (1, 2)]
    pf, served = _check_stream(PF.amazon_table_data_prefetcher, tb, lambda b: b, "cpu")
    for want, (field_value, label) in zip(tb, served):                                   # src/table_pretrain.py:246-
        assert len(field_value) == 6 and all(torch.equal(x, y) for x, y in zip(field_value, want[:6])) and torch.equal(label, want[6])
    cfg = ModelConfig(dataset="yelp", **SMALL)
    tb = [tuple(make_batch(cfg, 2, seed=s, n_reviews=2).field_value) + (torch.randint(0, 50, (2, 16), generator=g),) for s in (1, 2)]
    pf, served = _check_stream(PF.yelp_table_data_prefetcher, tb, lambda b: b, "cpu")
    assert all(torch.equal(x, y) for x, y in zip(served[1][0], tb[1][:6]))
    tx = [(b.reviews, b.reviews_mask, b.reviews_rating) for b in _loader(cfg, [2, 2])]
    pf, served = _check_stream(PF.text_data_prefetcher, tx, lambda b: b, "cpu")
    assert torch.equal(served[1][0], tx[1][0]) and served[1][1].max_review_len == int(tx[1][1].sum(-1).max())
    assert [o[0].shape[0] for o in PF.text_data_prefetcher(tx, device="cpu")] == [2, 2]   # iteration sugar


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [True, False])
def test_prefetcher_contents_under_overlap_gpu(pinned):
    """Copies of batch i+1 overlap work on batch i; slots are re-used every 2 batches (n_stage=2, the tightest ring) while the
    consumer stream is kept busy, so a slot rewritten too early or a batch handed out before its copy finished shows up as a
    content mismatch."""
    cfg = ModelConfig(dataset="yelp", **SMALL)
    batches = _loader(cfg, [8] * 7 + [3])
    tf = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
    pf = PF.yelp_data_prefetcher([tuple(tf(t) for t in _yelp_tuple(b)) for b in batches], n_stage=2)
    big = torch.randn(4096, 4096, device="cuda")
    sums = []
    out = pf.next()
    n = 0
    while out[0] is not None:
        for _ in range(3):
            big = torch.tanh(big @ big * 1e-3)                   # keeps the consumer stream tens of ms behind the host
        sums.append([t.clone() for t in _flat(out)])             # reads the slot on the consumer stream, after the queued work
        assert out[1].max_review_len == int(batches[n].reviews_mask.sum(-1).max())
        out = pf.next()
        n += 1
    torch.cuda.synchronize()
    assert n == len(batches)
    for b, got in zip(batches, sums):
        for g_, want in zip(got, _yelp_tuple(b)):
            assert g_.is_cuda and torch.equal(g_.cpu(), want)
    assert pf.allocations == 11 * 2


@pytest.mark.gpu
def test_reference_training_loop_body_through_the_prefetcher_gpu():
    """The loop of src/multimodal_train.py:353-379 verbatim (prefetcher.next() -> model(...) -> zero_grad / backward) on the drop-in:
    losses equal those of direct calls on the same batches, and the encoder ran on frames trimmed by the host-side hint (no mask
    read-back)."""
    from multimodalsum_b200.modules import MultimodalSum, YelpTableEncoder
    cfg = ModelConfig(dataset="yelp", **SMALL)
    sd = make_state_dict(cfg, seed=0, gates_open=True)
    batches = _loader(cfg, [2, 2, 2])

    def fresh():
        m = MultimodalSum(TableEncoder=YelpTableEncoder, config=cfg, label_smoothing=0.1)
        m.load_state_dict(sd, strict=False)
        return m.cuda().train()

    model = fresh()
    field = batches[0].field.cuda()
    want = []
    for b in batches:
        d = b.to("cuda")
        loss = model(d.reviews, d.reviews_mask, d.reviews_rating, field, d.field_value, d.img, d.img_mask)[0]
        model.zero_grad()
        loss.backward()
        want.append(loss.item())

    model = fresh()
    prefetcher = PF.yelp_data_prefetcher([tuple(t.pin_memory() for t in _yelp_tuple(b)) for b in batches])
    got = []
    reviews, reviews_mask, reviews_rating, field_value, img, img_mask = prefetcher.next()
    while reviews is not None:
        loss = model(reviews, reviews_mask, reviews_rating, field, field_value, img, img_mask)[0]
        model.zero_grad()
        loss.backward()
        got.append(loss.item())
        reviews, reviews_mask, reviews_rating, field_value, img, img_mask = prefetcher.next()
    assert got == want                                           # same kernels, same inputs, dropout 0: bit-equal losses
    assert model.engine._len_cache is None                       # the hint was used: the mask was never read back


def test_synthetic_dataset_through_dataloader_and_prefetcher_cpu():
    """The reference's data path end to end on the host: Dataset.__getitem__ (one business, src/multimodal_train.py:63-86) ->
    DataLoader collate (drop_last=False) -> prefetcher.next() tuples of the shapes MultimodalSum.forward takes."""
    from torch.utils.data import DataLoader
    from multimodalsum_b200.synth import SyntheticDataset
    for dataset, n_tab, n_img in (("yelp", 47, 10), ("amazon", 6, 1)):
        cfg = ModelConfig(dataset=dataset, **SMALL)
        data = SyntheticDataset(cfg, 5, seed=3, n_reviews=3, fixed_len=40)
        assert len(data) == 5 and data.field.shape[0] == n_tab and len(data[0]) == 11
        assert all(torch.equal(a, b) for a, b in zip(data[2], data[2]))                 # deterministic items
        first = data[0][0].clone()
        loader = DataLoader(data, batch_size=2, shuffle=False, num_workers=0, drop_last=False)
        cls = PF.yelp_data_prefetcher if dataset == "yelp" else PF.amazon_data_prefetcher
        sizes = []
        for reviews, reviews_mask, reviews_rating, field_value, img, img_mask in cls(loader, device="cpu"):
            B = reviews.shape[0]
            sizes.append(B)
            assert reviews.shape == (B, 3, 128) and reviews_mask.shape == (B, 3, 128) and reviews_rating.shape == (B, 3)
            assert img.shape == (B, n_img, 196, 1024) and img_mask.shape == (B, n_img) and img_mask.dtype == torch.bool
            assert len(field_value) == 6 and all(v.shape[0] == B and v.dtype == torch.int64 for v in field_value)
            assert reviews_mask.max_review_len == 40
        assert sizes == [2, 2, 1]
        data.set_epoch()
        assert not torch.equal(data[0][0], first)                                       # a new epoch draws new items
        with pytest.raises(IndexError):
            data[5]
