for i in 1 2 3 4 5; do
  python -m pytest tests/test_generate_gpu.py "tests/test_kernels_gpu.py::test_kernels_ignore_stale_onchip_state" "tests/test_step_gpu.py::test_graph_replayed_step_matches_plain_launches" -q -x 2>&1 | grep -v "^  *+\|^E  *+" | tail -30 | cut -c1-500
done
