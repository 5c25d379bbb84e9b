set -x
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step_final.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1
tail -2 gpurun_out/prof_step.log
MMSUM_NCU=1 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 6 -f -o gpurun_out/r02_gemm_2sm python tools/profile_gemm_cluster.py > gpurun_out/prof_gemm.log 2>&1
tail -2 gpurun_out/prof_gemm.log
ncu --set full --clock-control none --import-source on -k regex:attn_ --launch-skip 12 --launch-count 3 -f -o gpurun_out/r02_attn_final python tools/gpu_bench_attn.py > gpurun_out/prof_attn.log 2>&1
tail -2 gpurun_out/prof_attn.log
ls -la gpurun_out/*.ncu-rep
