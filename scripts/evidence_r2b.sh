# round-2 final evidence on ONE B200: bench line + GEMM table, ncu launch lists (time + DRAM bytes per launch) of one
# Stage-II / dense / Stage-I step, ncu --set full of the kernels added in the second half of round 2
set -x
cd $GRAFT_REPO_ROOT
ACT_BENCH_GEMM_TABLE=gpurun_out/r2_gemm_table_b.json timeout 600 python bench.py > gpurun_out/r2_bench_1gpu_c.json 2> gpurun_out/r2_bench_1gpu_c.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/l_step.csv python scripts/profile_step.py 1 128 native > gpurun_out/ncu_step.log 2>&1
python scripts/summarize_launches.py gpurun_out/l_step.csv gpurun_out/r2_gemm_traffic.json > gpurun_out/r2_launches_step_ncu.txt 2>&1
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/l_dense.csv python scripts/profile_step.py 1 16 synthetic 8192 512 > gpurun_out/ncu_dense.log 2>&1
python scripts/summarize_launches.py gpurun_out/l_dense.csv > gpurun_out/r2_launches_dense_ncu.txt 2>&1
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/l_dvae.csv python scripts/profile_dvae.py 64 > gpurun_out/ncu_dvae.log 2>&1
python scripts/summarize_launches.py gpurun_out/l_dvae.csv > gpurun_out/r2_launches_dvae_ncu.txt 2>&1
rm -f gpurun_out/l_step.csv gpurun_out/l_dense.csv gpurun_out/l_dvae.csv
timeout 400 ncu --profile-from-start off --set full --clock-control none -k regex:"gumbel_softmax|softmax_colmean|fold_input|kl_uniform" -c 8 -f -o gpurun_out/r2_stage1_new python scripts/profile_dvae.py 64 > gpurun_out/ncu_s1.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_stage1_new.ncu-rep > gpurun_out/r2_ncu_stage1_kernels_full.txt 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none -k regex:"fps_cluster|knn_kernel" -c 2 -f -o gpurun_out/r2_fps_cluster python scripts/profile_step.py 1 16 synthetic 8192 512 > gpurun_out/ncu_fps.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_fps_cluster.ncu-rep > gpurun_out/r2_ncu_dense_tokenizer_full.txt 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none -k regex:"gemm_bf16_pair_kernel" -c 6 -f -o gpurun_out/r2_gemm_pair python scripts/profile_step.py 1 128 native > gpurun_out/ncu_pair.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_gemm_pair.ncu-rep > gpurun_out/r2_ncu_gemm_pair_full.txt 2>&1
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -20
head -30 gpurun_out/r2_launches_step_ncu.txt
