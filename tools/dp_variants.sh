#!/bin/bash
# N-GPU bench under different gradient all-reduce settings (one gpurun --gpus N call): wire dtype x NCCL knobs.
N=${1:-8}
run() {   # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_dp_${N}_${name}.json 2> gpurun_out/r02_dp_${N}_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_dp_${N}_${name}.json").read().strip().splitlines()[-1])
    g = d["roofline_kernels"]["gemm"]
    print("%-22s value %8.1f  ms/step %7.2f  gemm avg %6.1f us frac %.3f  e2e %8.1f" % ("${name}", d["value"], d["ms_per_step"], g["avg_us"], g["frac"], d["e2e"]["value"]))
except Exception as e:
    print("${name} FAILED", e)
PY
}
run bf16_default MMSUM_DP_DTYPE=bf16
run fp32_default MMSUM_DP_DTYPE=fp32
run fp32_nvls MMSUM_DP_DTYPE=fp32 NCCL_ALGO=NVLS
run bf16_maxcta4 MMSUM_DP_DTYPE=bf16 NCCL_MAX_CTAS=4
run fp32_maxcta4 MMSUM_DP_DTYPE=fp32 NCCL_MAX_CTAS=4
run fp32_maxcta16 MMSUM_DP_DTYPE=fp32 NCCL_MAX_CTAS=16
timeout 200 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02_dp_${N}_n1.json 2>/dev/null
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_dp_${N}_n1.json").read().strip().splitlines()[-1]); g = d["roofline_kernels"]["gemm"]
print("%-22s value %8.1f  ms/step %7.2f  gemm avg %6.1f us frac %.3f" % ("N=1 same box", d["value"], d["ms_per_step"], g["avg_us"], g["frac"]))
PY
