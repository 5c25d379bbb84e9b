#!/bin/bash
# Weak-scaling sweep on one box: tools/dp_scale.sh N  (prints businesses/s at 1 and N GPUs; torchrun, NCCL)
N=${1:-2}
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   n=%d  %.2f businesses/s  %.2f ms/step  e2e %.2f  gemm frac %.3f  clk %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks']['sm_mhz']))"; }
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | show
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>gpurun_out/dp_scale.err | show
