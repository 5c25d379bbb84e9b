timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | tail -3
for L in as ""; do
  if [ -z "$L" ]; then unset MMSUM_LIB_PATH; else export MMSUM_LIB_PATH=$PWD/multimodalsum_b200/libmmsum_b200_$L.so; fi
  echo "=== lib '$L'"; timeout 200 python tools/gpu_gemm_fc1.py
done
unset MMSUM_LIB_PATH
echo "=== split sweep (2SM)"; timeout 400 python tools/gpu_gemm_sweep.py
