timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention or stale" 2>&1 | tail -3
for L in base ""; do
  if [ -z "$L" ]; then unset MMSUM_LIB_PATH; else export MMSUM_LIB_PATH=$PWD/multimodalsum_b200/libmmsum_b200_$L.so; fi
  echo "=== lib '$L'"
  MMSUM_ATTN_BWD_PART=1 timeout 120 python tools/gpu_bench_attn.py | grep bwd | sed 's/^/dQ  /'
  MMSUM_ATTN_BWD_PART=2 timeout 120 python tools/gpu_bench_attn.py | grep bwd | sed 's/^/dKV /'
done
