#!/bin/bash
# Same-box A/B of step-level switches: tools/ab_step.sh "NAME=VAL" "NAME2=VAL2" ...  (each run: default vs that setting, interleaved)
# Prints businesses/s and ms/step per run; box-to-box variance (+-3 %) makes cross-run comparisons meaningless.
run() { env "$@" python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   %.2f businesses/s  %.2f ms/step  e2e %.2f  gemm %.1f us  clk %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_kernels']['gemm']['avg_us'], d['clocks']['sm_mhz']))"; }
for rep in 1 2; do
  echo "default (rep $rep)"; run X=1
  for kv in "$@"; do echo "$kv (rep $rep)"; run $kv; done
done
