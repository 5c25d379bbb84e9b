for L in "" scal; do
  if [ -z "$L" ]; then unset MMSUM_LIB_PATH; else export MMSUM_LIB_PATH=$PWD/multimodalsum_b200/libmmsum_b200_$L.so; fi
  echo "=== lib '$L'"
  MMSUM_ATTN_BWD_PART=1 timeout 120 python tools/gpu_bench_attn.py | sed 's/^/dQ  /'
  MMSUM_ATTN_BWD_PART=2 timeout 120 python tools/gpu_bench_attn.py | grep bwd | sed 's/^/dKV /'
done
for L in "" gscal; do
  if [ -z "$L" ]; then unset MMSUM_LIB_PATH; else export MMSUM_LIB_PATH=$PWD/multimodalsum_b200/libmmsum_b200_$L.so; fi
  echo "=== lib '$L'"; N_ITER=20 timeout 200 python tools/gpu_gemm_fc1.py
done
