#!/usr/bin/env python
"""bench.py — headline benchmark: MultimodalSum data-parallel training step, businesses/s (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]      the reference algorithm's CPU implementation (oracle port)
                                                                 on the host cores, same metric / config

One step = forward + backward (+ bucketed gradient all-reduce when N > 1) of `MultimodalSum.forward` over one batch of
synthetic Yelp-shaped businesses (BASELINE.json configs[1]: 16 businesses / GPU, 9 reviews x 128-token frame with 100
valid tokens, 47 table fields, 10 images x 196 pooled ResNet features, BART-large, bf16 compute, dropout 0.1).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TFLOP_PER_BUSINESS = 4.187   # algorithmic fwd+bwd FLOPs per Yelp business, SURVEY.md §8(d) / App. C (Amazon: 3.640)
METRIC = "train businesses/sec (BART-large, Yelp shape)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--businesses", type=int, default=16, help="businesses per GPU (BASELINE config 2: 16)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dataset", default="yelp", choices=["yelp", "amazon"],
                    help="yelp = BASELINE configs[1-2] (headline); amazon = configs[3] shape (133 table fields, 1 image, 70 valid tokens)")
    ap.add_argument("--cpu-seconds", type=float, default=240.0, help="time budget of the reference arm")
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], source="measured")
    except Exception:
        return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm / cpu baseline
def cpu_reference_steps(steps, warmup, budget_s):
    """Time the reference algorithm's CPU implementation (oracle port of MultimodalSum.forward + backward, as-written
    9-pass loop, fp32, dropout 0.1, no optimizer — BASELINE.md §3) on the host cores.  One step = ONE business."""
    import torch
    from multimodalsum_b200.synth import ModelConfig, make_batch, make_state_dict
    from oracle import mmsum_oracle as OR
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ModelConfig(dataset="yelp", dropout=0.1)
    sd = make_state_dict(cfg, seed=0, perturb=False)
    batch = make_batch(cfg, 1, seed=1, fixed_len=100, n_valid_imgs=10)
    times = []
    t_start = time.time()
    n_done = 0
    for i in range(warmup + steps):
        t0 = time.time()
        OR.step_loss_and_grads(sd, cfg, batch, 0.1, dtype=torch.float32, device="cpu", training=True)
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        n_done += 1
        elapsed = time.time() - t_start
        if i < warmup and elapsed + dt * (warmup - i - 1 + steps) > budget_s:
            warmup = i + 1                      # cut warm-up short to keep the run within the budget
        if times and elapsed + dt > budget_s:
            break
    if not times:
        times = [dt]
    med = statistics.median(times)
    return dict(value=1.0 / med, ms_per_step=1000.0 * med, steps=len(times), warmup=min(warmup, n_done - len(times)), cores=cores,
                sample="1 business per step (B=1, full Yelp shape: 9x128-token reviews, 47 table fields, 10x196 image keys), "
                       "oracle port fwd+bwd, fp32, dropout 0.1, %d timed step(s), median" % len(times))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_steps(args.steps, args.warmup, args.cpu_seconds)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "businesses/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # same workload as the GPU arm (BASELINE configs[1] business shape); each timed step is a bounded sample of it: one business
        "config": {"workload": "BASELINE configs[1]: Yelp-shape multimodal_train step  (fwd + bwd), BART-large random init",
                   "businesses_per_gpu": 16, "reviews": 9, "frame": 128, "valid_tokens": 100, "table_fields": 47, "images": "10x196",
                   "dropout": 0.1, "label_smoothing": 0.1, "parallelism": "cpu x%d threads" % r["cores"],
                   "sample": "1 business per timed step (the reference's per-business cost is independent of the batch)"},
        "cpu_baseline": {"value": r["value"], "unit": "businesses/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "businesses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ this repo's arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    from multimodalsum_b200 import ops
    from multimodalsum_b200.dp import GradAllReducer
    from multimodalsum_b200.modules import AmazonTableEncoder, MultimodalSum, YelpTableEncoder
    from multimodalsum_b200.optim import get_optimizer
    from multimodalsum_b200.synth import ModelConfig, make_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", init_method="env://", device_id=dev)
    B = args.businesses
    amazon = args.dataset == "amazon"
    cfg = ModelConfig(dataset=args.dataset, dropout=0.1)
    torch.manual_seed(0)                                   # identical random-init bart-large weights on every rank
    model = MultimodalSum(TableEncoder=AmazonTableEncoder if amazon else YelpTableEncoder, config=cfg, label_smoothing=0.1).to(dev).train()
    # distinct synthetic shards per rank (businesses are independent units: weak scaling, no data-path collective)
    n_host_batches = 2
    host = [make_batch(cfg, B, seed=1234 + rank * 17 + i, fixed_len=70 if amazon else 100, n_valid_imgs=1 if amazon else 10).pin()
            for i in range(n_host_batches)]
    resident = host[0].to(dev)
    h2d_bytes = host[0].nbytes()

    opt = [None]

    def step(batch):
        # src/multimodal_train.py:357-364: forward, zero_grad, backward (+ bucketed all-reduce), clip, AdamW, schedule
        loss = model(batch.reviews, batch.reviews_mask, batch.reviews_rating, batch.field, batch.field_value, batch.img, batch.img_mask)[0]
        model.zero_grad(set_to_none=True)
        loss.backward()
        if opt[0] is not None:
            opt[0].step(lr=1e-5)
        return loss

    step(resident)                                          # builds the engine, arenas and workspaces
    eng = model.engine
    reducer = GradAllReducer(eng) if world > 1 else None
    # all parameters in a group (a list, not the reference's exhausted generator: every tensor is updated here)
    opt[0] = get_optimizer(eng, 1e-5, ["bias", "LayerNorm.weight"], list(model.named_parameters()), None, max_grad_norm=1.0)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing (value) + live GEMM roofline
    for _ in range(max(args.warmup, 3)):
        step(resident)
    barrier()
    gemm_events = []

    @contextlib.contextmanager
    def gemm_timer(flops):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        yield
        e.record()
        gemm_events.append((flops, s, e))

    ops.GEMM_TIMER = gemm_timer
    launches0 = ops.LAUNCHES
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for _ in range(args.steps):
        step(resident)
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ops.GEMM_TIMER = None
    launches = ops.LAUNCHES - launches0
    ms = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step / 1000.0)
    gemm_flops = sum(f for f, _, _ in gemm_events)
    gemm_ms = sum(s.elapsed_time(e) for _, s, e in gemm_events)
    n_gemm = len(gemm_events)

    # ---------------- end-to-end timing through the public API with HOST (pinned) inputs
    copy_stream = torch.cuda.Stream()

    def stage(i):
        with torch.cuda.stream(copy_stream):
            b = host[i % n_host_batches].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return b, ev

    for i in range(2):
        b, ev = stage(i)
        torch.cuda.current_stream().wait_event(ev)
        step(b).item()
    barrier()
    t0 = time.perf_counter()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    nxt = stage(0)
    loss_host = torch.empty(1, pin_memory=True)
    for i in range(args.steps):
        b, ev = nxt
        torch.cuda.current_stream().wait_event(ev)
        if i + 1 < args.steps:
            nxt = stage(i + 1)                               # prefetch the next step's inputs on the copy stream
        loss = step(b)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)   # device -> host read of the step's result
    e_end.record()
    barrier()
    e2e_ms = e_start.elapsed_time(e_end)
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    e2e_value = B * world / (e2e_ms / args.steps / 1000.0)
    final_loss = float(loss_host.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    tfb = 3.640 if amazon else TFLOP_PER_BUSINESS
    achieved_tf = gemm_flops / (gemm_ms / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": "businesses/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": ("BASELINE configs[3] shape: Amazon-shape multimodal_train step" if amazon else "BASELINE configs[1]: Yelp-shape multimodal_train step") + "  (fwd + bwd%s + fused clip/AdamW), BART-large random init" % (
                       " + bucketed NCCL grad all-reduce overlapped with backward" if world > 1 else ""),
                   "businesses_per_gpu": B, "reviews": 9, "frame": 128, "valid_tokens": 70 if amazon else 100,
                   "table_fields": 133 if amazon else 47, "images": "1x196" if amazon else "10x196", "dropout": 0.1, "label_smoothing": 0.1, "parallelism": "dp%d" % world,
                   "l2_policy": "per-step working set (>25 GB activations + 2.8 GB weights) exceeds the 126 MB L2",
                   "final_loss": final_loss},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "businesses/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "tc_fraction_step": {"algorithmic_tflop_per_business": tfb,
                             "achieved_tflops_per_gpu": value / world * tfb,
                             "frac_of_sustained_peak": value / world * tfb / peaks["tf_sustained"]},
        "roofline": {"kernel": "gemm_tcgen05_kernel (all %d GEMM launches of the timed steps)" % n_gemm, "bound": "tensor",
                     "achieved": achieved_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": achieved_tf / peaks["tf_sustained"], "traffic": None, "peak_source": peaks["source"] + " (sustained bf16)",
                     "gemm_share_of_step": gemm_ms / ms if ms > 0 else None},
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_steps(1, 0, 60.0)
        line["cpu_baseline"] = {"value": r["value"], "unit": "businesses/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
