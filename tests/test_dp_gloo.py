"""Data-parallel gradient exchange on CPU: world_size 2 over gloo.  Exercises the bucketing / ordering logic of
dp.GradAllReducer (the NCCL path differs only in ReduceOp.AVG and the side stream) and reduce_tensor."""
import os
import socket
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from multimodalsum_b200.dp import GradAllReducer, reduce_tensor
    n = 1 << 16
    eng = types.SimpleNamespace(G32=torch.zeros(n), numel=n, device=torch.device("cpu"), grad_ready_hook=None)
    red = GradAllReducer(eng, bucket_mb=0.0625)           # 16 Ki fp32 elements per bucket
    assert eng.grad_ready_hook is not None
    torch.manual_seed(100 + rank)
    local = torch.randn(n)
    # backward reports completion layer by layer (arena order == completion order)
    cuts = [1000, 9000, 20000, 20001, 50000, n]
    lo = 0
    for hi in cuts:
        eng.G32[lo:hi] = local[lo:hi]
        eng.grad_ready_hook(hi)
        lo = hi
    # expected: the average over ranks
    torch.manual_seed(100)
    g0 = torch.randn(n)
    torch.manual_seed(101)
    g1 = torch.randn(n)
    ok = torch.allclose(eng.G32, (g0 + g1) / 2, atol=1e-6)
    # small ranges are merged into >= bucket-size all-reduces; the tail is always flushed
    ok = ok and red.n_buckets == 3 and red.lo == 0 and not red.works
    # DDP(model, delay_allreduce=True) of src/multimodal_train.py:474: rank 0's parameters reach every rank, the reducer is hooked
    from multimodalsum_b200.dp import DistributedDataParallel as DDP

    class _Model(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(4))
            self.eng = types.SimpleNamespace(W32=torch.full((64,), float(rank + 1)), G32=torch.zeros(64), numel=64, device=torch.device("cpu"),
                                             grad_ready_hook=None, dirty=0)
            self.eng.mark_weights_dirty = lambda: setattr(self.eng, "dirty", self.eng.dirty + 1)

        def _ensure_engine(self, device):
            return self.eng

        def forward(self, x, scale=1.0):
            return (x * scale,)

    m = _Model()
    ddp = DDP(m, delay_allreduce=True)
    ok = ok and ddp.module is m and bool((m.eng.W32 == 1.0).all()) and m.eng.dirty == 1 and m.eng.grad_ready_hook is not None
    ok = ok and ddp(torch.ones(1), scale=3.0)[0].item() == 3.0 and list(ddp.state_dict()) == ["module.w"]
    r = reduce_tensor(torch.tensor(float(rank + 1)), world)
    ok = ok and abs(r.item() - 1.5) < 1e-6
    q.put((rank, bool(ok), red.n_buckets))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
