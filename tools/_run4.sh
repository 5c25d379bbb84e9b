export MMSUM_GEMM_CLUSTER=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm" 2>&1 | grep -v "^  *+\|^E  *+" | tail -15 | cut -c1-300
echo "=== cluster on"; timeout 200 python tools/profile_gemm_cluster.py
export MMSUM_GEMM_CLUSTER=0
echo "=== cluster off"; timeout 200 python tools/profile_gemm_cluster.py
