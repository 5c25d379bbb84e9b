timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_step_gpu.py -q -x -k "attention or stale or uninit or graph_replayed or small" 2>&1 | tail -3
for L in base nodq ""; do
  if [ -z "$L" ]; then unset MMSUM_LIB_PATH; else export MMSUM_LIB_PATH=$PWD/multimodalsum_b200/libmmsum_b200_$L.so; fi
  echo "=== lib '$L'"
  MMSUM_ATTN_BWD_PART=1 timeout 120 python tools/gpu_bench_attn.py | sed 's/^/dQ  /'
done
