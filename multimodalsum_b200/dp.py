"""Data-parallel gradient exchange: bucketed all-reduce over NCCL (NVLink 5 / NVSwitch) overlapped with backward.

Reference behaviour replaced: apex `DistributedDataParallel(model, delay_allreduce=True)` (src/multimodal_train.py:474) —
flatten ALL gradients, one all-reduce after backward has finished, divide by world size; no overlap (quirk Q10).
Here the gradient arena is ordered by completion time (engine._arena_order), the engine reports "arena[0:hi) is
final" after every layer, and each bucket (a contiguous fp32 slice, no flatten / unflatten copies) is all-reduced
(AVG) on NCCL's stream while the compute stream keeps running backward.  One process per GPU; businesses are
independent, so there is no data-path collective other than this one (plus the scalar loss for logging,
src/utils.py:8-12).
"""
import torch
import torch.distributed as dist


class GradAllReducer:
    def __init__(self, engine, process_group=None, bucket_mb=64):
        self.engine = engine
        self.pg = process_group
        self.bucket_elems = int(bucket_mb * (1 << 20) // 4)
        self.lo = 0
        self.works = []
        self.n_buckets = 0
        self.comm_stream = torch.cuda.Stream() if engine.device.type == "cuda" else None
        engine.grad_ready_hook = self.on_ready

    def world(self):
        return dist.get_world_size(self.pg) if dist.is_initialized() else 1

    def on_ready(self, hi):
        eng = self.engine
        last = hi >= eng.numel
        if hi - self.lo < self.bucket_elems and not last:
            return
        if self.world() > 1:
            bucket = eng.G32[self.lo:hi]
            if self.comm_stream is not None:
                ev = torch.cuda.Event()
                ev.record()                               # gradients of this bucket are enqueued on the compute stream
                self.comm_stream.wait_event(ev)
                with torch.cuda.stream(self.comm_stream):
                    self.works.append(dist.all_reduce(bucket, op=dist.ReduceOp.AVG, group=self.pg, async_op=True))
            else:                                          # gloo (CPU tests): no AVG, no streams
                w = dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
                self.works.append((w, bucket))
        self.n_buckets += 1
        self.lo = hi
        if last:
            self.finish()

    def finish(self):
        """Make the compute stream wait for every outstanding bucket (no host sync on CUDA)."""
        for w in self.works:
            if isinstance(w, tuple):
                w[0].wait()
                w[1].div_(self.world())
            else:
                w.wait()
        self.works = []
        self.lo = 0


def reduce_tensor(tensor, world_size):
    """src/utils.py:8-12 — average a scalar over ranks for logging."""
    rt = tensor.clone()
    if dist.is_initialized() and world_size > 1:
        dist.all_reduce(rt, op=dist.ReduceOp.SUM)
    rt /= world_size
    return rt
